#!/bin/bash
# Round-2 profiling visit: ncu launch list of the default bench command, launch lists + `--set full` captures of the ranking path
# at C3 / T1M (+ C5 launch list), per-line stall samples of the hot kernels, DRAM traffic per launch. usage: tools/gpu_ncu_r2.sh <tag>
TAG=${1:-r2n1}
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_bench_default_$TAG.csv python bench.py --steps 2 --warmup 1 --no-sharded --no-cpu-baseline --no-extra > $O/ncu_bench_default_$TAG.log 2>&1
python tools/summarize_launches.py $O/launches_bench_default_$TAG.csv > $O/launches_bench_default_$TAG.txt 2>&1
KEEP='regex:exact_|gram|screen|press_chk|xb_kernel|zscore|weights_dmma|col_chunk|select|pls_defl|eref|pack_|decide|pls_u|pls_r|dist_scores|wide_'
for W in ${2:-C3 T1M}; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_${W}_$TAG.csv python tools/profile_rank.py $W 2 > $O/ncu_l_${W}_$TAG.log 2>&1
  python tools/summarize_launches.py $O/launches_${W}_$TAG.csv > $O/launches_${W}_$TAG.txt 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k "$KEEP" --profile-from-start off -c 90 -f -o /tmp/full_${W}_$TAG python tools/profile_rank.py $W 2 > $O/ncu_f_${W}_$TAG.log 2>&1
  ncu -i /tmp/full_${W}_$TAG.ncu-rep --page raw --csv > $O/full_${W}_$TAG.raw.csv 2>/dev/null
  python tools/ncu_brief.py /tmp/full_${W}_$TAG.ncu-rep > $O/full_${W}_${TAG}_brief.txt 2>&1
  for KN in exact_rank_kernel exact_scatter_kernel screen1_kernel screen2_kernel press_chk_kernel xb_kernel pls_defl_kernel gram_kernel; do
    python tools/ncu_lines.py /tmp/full_${W}_$TAG.ncu-rep $KN 30 > $O/lines_${W}_${KN}_$TAG.txt 2>&1
  done
  python tools/ncu_traffic.py $O/full_${W}_$TAG.raw.csv > $O/traffic_${W}_$TAG.json 2>/dev/null
  gzip -f $O/full_${W}_$TAG.raw.csv
done
head -14 $O/launches_bench_default_$TAG.txt; du -sh $O
