#!/bin/bash
# Short GPU-box visit: parity tests, PLS phase profile of both component loops, bench lines. usage: tools/gpu_quick.sh <tag>
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
for W in C2 C3; do
  ABCB200_PLS_PROF=1 timeout 300 python tools/profile_rank.py $W 2 > $O/prof_${W}_$TAG.txt 2>&1
  ABCB200_PLS_LITERAL=1 ABCB200_PLS_PROF=1 timeout 300 python tools/profile_rank.py $W 2 > $O/prof_${W}_literal_$TAG.txt 2>&1
done
timeout 600 python bench.py > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err
timeout 600 python bench.py --workload C3 --steps 10 --warmup 5 --no-sharded > $O/bench_c3_$TAG.json 2> $O/bench_c3_$TAG.err
tail -5 $O/pytest_gpu_$TAG.log; cut -c1-400 $O/bench_c2_$TAG.json; grep "pls_" $O/prof_*_$TAG.txt | cut -c1-400
