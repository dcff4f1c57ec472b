#!/bin/bash
# GPU-box visit: parity tests only (optionally a -k filter). usage: tools/gpu_tests.sh <tag> [pytest -k expression]
TAG=${1:-t}
O=gpurun_out
mkdir -p $O
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $O/pytest_gpu_$TAG.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1
fi
echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -40 $O/pytest_gpu_$TAG.log
