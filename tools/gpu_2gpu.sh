#!/bin/bash
# 2-GPU visit: the C++ host sharded over both GPUs, then the bench under torchrun with 2 ranks (C-side NCCL group)
TAG=${1:-g2}; O=gpurun_out; mkdir -p $O
nvidia-smi -L
g++ -std=c++17 -O1 tests/cpp/sharded_test.cpp -Labcsmc_b200 -labcsmc_b200 -Wl,-rpath,$PWD/abcsmc_b200 -o /tmp/sharded_test && \
  for G in 1 2; do timeout 300 /tmp/sharded_test $G 3001 2000 30; echo "rc=$?"; done 2>&1 | tee $O/sharded_test_$TAG.log
timeout 300 /tmp/sharded_test 2 200000 100000 30 2>&1 | tee -a $O/sharded_test_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > $O/bench_default_${TAG}.json 2> $O/bench_default_${TAG}.err
tail -3 $O/bench_default_${TAG}.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_default_g2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","e2e")}); print(d["sharded_weight_update"])
P
