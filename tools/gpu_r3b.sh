#!/bin/bash
# visit: GPU tests, then A/B bench lines of the exact level (fine bins vs radix)
TAG=${1:-r3b}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|Error|assert" $O/pytest_gpu_$TAG.log | head -30 | cut -c1-300
bash tools/gpu_ab.sh $TAG "${2:-C3 T1M}" "${3:-base radix=ABCB200_EXACT_RADIX=1}"
