#!/bin/bash
# N-GPU visit: the C++ host sharded over all GPUs, then the bench under torchrun with N ranks
N=${1:-8}; TAG=${2:-g$N}; O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l
g++ -std=c++17 -O1 tests/cpp/sharded_test.cpp -Labcsmc_b200 -labcsmc_b200 -Wl,-rpath,$PWD/abcsmc_b200 -o /tmp/sharded_test && timeout 300 /tmp/sharded_test $N 400000 100000 30 2>&1 | tee $O/sharded_test_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_default_${TAG}.json 2> $O/bench_default_${TAG}.err
tail -3 $O/bench_default_${TAG}.err | cut -c1-300; python - <<P
import json
d=json.loads(open("$O/bench_default_${TAG}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["ms_per_step"]); s=d["sharded_weight_update"]; print({k:s[k] for k in ("ms_per_step","value","max_rel_err","max_rel_err_vs_oracle")}, s["roofline"]["frac"])
P
