#!/bin/bash
# Quick iteration visit: parity tests, PLS phase cycles, then C2 / C3 (/ C5 with a third argument) bench lines without the CPU leg.
# usage: tools/gpu_iter.sh <tag> [pytest -k expr | all] [c5]
TAG=${1:-it}
O=gpurun_out
mkdir -p $O
if [ -n "$2" ] && [ "$2" != "all" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $O/pytest_gpu_$TAG.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1
fi
echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
for W in C2 C3; do ABCB200_PLS_PROF=1 timeout 300 python tools/profile_rank.py $W 2 > $O/prof_${W}_$TAG.txt 2>&1; done
timeout 300 python bench.py --no-sharded --no-cpu-baseline > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err
timeout 300 python bench.py --workload C3 --steps 10 --warmup 5 --no-sharded --no-cpu-baseline > $O/bench_c3_$TAG.json 2> $O/bench_c3_$TAG.err
if [ -n "$3" ]; then
  timeout 600 python bench.py --workload C5 --steps 3 --warmup 3 --no-sharded --no-cpu-baseline > $O/bench_c5_$TAG.json 2> $O/bench_c5_$TAG.err
fi
tail -4 $O/pytest_gpu_$TAG.log
grep -h "pls_" $O/prof_C2_$TAG.txt $O/prof_C3_$TAG.txt | sort -u | cut -c1-300
python tools/bench_brief.py $O/bench_c2_$TAG.json $O/bench_c3_$TAG.json $O/bench_c5_$TAG.json 2>/dev/null
