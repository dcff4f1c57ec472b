#!/bin/bash
# launch lists (ncu gpu__time_duration) of one ranking pass: tools/gpu_launchlist.sh TAG "C3 T1M"
TAG=${1:-ll}; O=gpurun_out; mkdir -p $O
for W in ${2:-C3 T1M}; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_${W}_$TAG.csv python tools/profile_rank.py $W 2 > $O/ncu_l_${W}_$TAG.log 2>&1
  python tools/summarize_launches.py $O/launches_${W}_$TAG.csv > $O/launches_${W}_$TAG.txt 2>&1
  head -30 $O/launches_${W}_$TAG.txt | cut -c1-110
done
