#!/bin/bash
TAG=${1:-r3d}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x -k "selection or exact or golden or wilcoxon or ranking_pls" > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|Error|assert" $O/pytest_gpu_$TAG.log | head -30 | cut -c1-300
bash tools/gpu_ab.sh $TAG "${2:-C3 T1M}" "${3:-base}"
bash tools/gpu_launchlist.sh $TAG "T1M" | grep -E "exact|screen2|radix|launches"
