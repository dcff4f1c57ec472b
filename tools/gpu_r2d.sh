#!/bin/bash
# round-2 visit D: pipelined fit + hold-out (green-context lanes): parity tests, then A/B bench of C3 / T1M / C2 with and without it
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -4 $O/pytest_gpu_$TAG.log
for W in C3 T1M C2; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-sharded --no-cpu-baseline --no-extra > $O/bench_${W}_$TAG.json 2> $O/bench_${W}_$TAG.err
  ABCB200_NO_PIPELINE=1 timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-sharded --no-cpu-baseline --no-extra > $O/bench_${W}_${TAG}_nopipe.json 2> $O/bench_${W}_${TAG}_nopipe.err
  ABCB200_NO_GREEN=1 timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-sharded --no-cpu-baseline --no-extra > $O/bench_${W}_${TAG}_nogreen.json 2> $O/bench_${W}_${TAG}_nogreen.err
  for v in "" _nopipe _nogreen; do python tools/bench_brief.py $O/bench_${W}_${TAG}$v.json 2>/dev/null | head -14; tail -2 $O/bench_${W}_${TAG}$v.err; done
done
