#!/bin/bash
# GPU-box visit: smoke, parity tests, default bench line (+ reference arm), C3 bench line. usage: tools/gpu_v8.sh <tag>
TAG=${1:-v8}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 600 python bench.py > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c2_$TAG.json 2> $O/bench_ref_c2_$TAG.err
timeout 600 python bench.py --workload C3 --steps 10 --warmup 5 --no-sharded > $O/bench_c3_$TAG.json 2> $O/bench_c3_$TAG.err
tail -3 $O/smoke_$TAG.log; tail -15 $O/pytest_gpu_$TAG.log; cut -c1-300 $O/bench_c2_$TAG.json; cut -c1-300 $O/bench_ref_c2_$TAG.json; cut -c1-300 $O/bench_c3_$TAG.json
