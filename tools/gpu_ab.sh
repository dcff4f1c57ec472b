#!/bin/bash
# A/B bench lines: tools/gpu_ab.sh TAG "W1 W2" "VARIANT=ENV ..."   (variant "base" = no extra environment)
TAG=$1; O=gpurun_out; mkdir -p $O
for W in $2; do
  for V in $3; do
    name=${V%%=*}; envs=${V#*=}; [ "$name" = "$V" ] && envs=""; envs=${envs//:/ }
    env ABCB200_DEBUG=1 $envs timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-sharded --no-cpu-baseline --no-extra > $O/bench_${W}_${TAG}_$name.json 2> $O/bench_${W}_${TAG}_$name.err
    python tools/bench_brief.py $O/bench_${W}_${TAG}_$name.json > $O/brief_${W}_${TAG}_$name.txt 2>/dev/null; head -3 $O/brief_${W}_${TAG}_$name.txt | cut -c1-400; grep -o "'sm_partition': [01]" $O/brief_${W}_${TAG}_$name.txt; tail -2 $O/bench_${W}_${TAG}_$name.err
  done
done
