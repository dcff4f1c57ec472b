#!/bin/bash
TAG=${1:-r2g}; O=gpurun_out; mkdir -p $O
python tools/h2d_probe.py 2>&1 | tee $O/h2d_probe_$TAG.txt
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -30 $O/pytest_gpu_$TAG.log | cut -c1-300
bash tools/gpu_ab.sh $TAG "C5" "base nopipe=ABCB200_NO_PIPELINE=1"
