#!/bin/bash
TAG=${1:-r2i}; O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu_$TAG.log | head -30 | cut -c1-300
bash tools/gpu_ab.sh $TAG "C3 T1M C2 C5" "base"
for W in C3 T1M C5; do grep -E "screen1|screen2" $O/brief_${W}_${TAG}_base.txt; done
