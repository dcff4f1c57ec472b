#!/bin/bash
TAG=${1:-t}
O=gpurun_out
mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q ${2:-} > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -25 $O/pytest_gpu_$TAG.log
