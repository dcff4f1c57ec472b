#!/bin/bash
# GPU-box visit: parity tests, C5 bench line, ncu launch list of the default bench command. usage: tools/gpu_v7.sh <tag>
TAG=${1:-v7}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt
free -g > $O/mem_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_c2_$TAG.csv python bench.py --steps 2 --warmup 3 --no-sharded --no-cpu-baseline > $O/ncu_bench_c2_$TAG.log 2>&1
python tools/summarize_launches.py $O/launches_bench_c2_$TAG.csv > $O/launches_bench_c2_$TAG.txt 2>&1
timeout 900 python bench.py --workload C5 --steps 3 --warmup 3 --no-sharded > $O/bench_c5_$TAG.json 2> $O/bench_c5_$TAG.err
tail -3 $O/pytest_gpu_$TAG.log; cut -c1-600 $O/bench_c5_$TAG.json; tail -3 $O/bench_c5_$TAG.err
