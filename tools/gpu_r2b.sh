#!/bin/bash
# round-2 visit B: GPU parity tests + the default bench line (C3 + configs + sharded C4)
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 900 python bench.py > $O/bench_default_$TAG.json 2> $O/bench_default_$TAG.err
tail -15 $O/pytest_gpu_$TAG.log; cut -c1-600 $O/bench_default_$TAG.json; tail -5 $O/bench_default_$TAG.err
