#!/bin/bash
# round-2 visit A: GPU parity tests (incl. full-size goldens) + baseline bench lines for C3 / T1M
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 600 python bench.py --workload C3 --steps 10 --warmup 5 --no-sharded --no-cpu-baseline > $O/bench_C3_$TAG.json 2> $O/bench_C3_$TAG.err
timeout 600 python bench.py --workload T1M --steps 5 --warmup 3 --no-sharded --no-cpu-baseline > $O/bench_T1M_$TAG.json 2> $O/bench_T1M_$TAG.err
tail -15 $O/pytest_gpu_$TAG.log; cut -c1-400 $O/bench_C3_$TAG.json; cut -c1-400 $O/bench_T1M_$TAG.json
