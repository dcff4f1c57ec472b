#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch lists and full captures. Outputs under gpurun_out/ (<= 64 MiB:
# the .ncu-rep files are summarised to CSV on the box and deleted).
# usage: tools/gpu_round.sh <tag> [skip_full]
TAG=${1:-v3}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt
./tools/lat_probe > $O/lat_probe_$TAG.txt 2>&1; ./tools/lat_probe2 > $O/lat_probe2_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 600 python bench.py > $O/bench_c2_$TAG.json 2> $O/bench_c2_$TAG.err
timeout 600 python bench.py --workload C3 --steps 10 --warmup 5 --no-sharded > $O/bench_c3_$TAG.json 2> $O/bench_c3_$TAG.err
KEEP='regex:gram|screen|press_chk|xb_kernel|zscore|weights_dmma|col_chunk|select|bitonic'
for W in C3 C2; do
  ABCB200_PLS_PROF=1 timeout 300 python tools/profile_rank.py $W 2 > $O/prof_${W}_$TAG.txt 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_${W}_$TAG.csv python tools/profile_rank.py $W 2 > $O/ncu_l_${W}_$TAG.log 2>&1
  python tools/summarize_launches.py $O/launches_${W}_$TAG.csv > $O/launches_${W}_$TAG.txt 2>&1
  if [ -z "$2" ]; then
    timeout 900 ncu --set full --clock-control none --import-source on -k "$KEEP" --profile-from-start off -c 40 -f -o /tmp/full_${W}_$TAG python tools/profile_rank.py $W 2 > $O/ncu_f_${W}_$TAG.log 2>&1
    ncu -i /tmp/full_${W}_$TAG.ncu-rep --page raw --csv > $O/full_${W}_$TAG.raw.csv 2>/dev/null
    python tools/ncu_brief.py /tmp/full_${W}_$TAG.ncu-rep > $O/full_${W}_${TAG}_brief.txt 2>&1
    for KN in gram_kernel pls_gram_kernel screen1_kernel press_chk_kernel xb_kernel; do
      python tools/ncu_lines.py /tmp/full_${W}_$TAG.ncu-rep $KN 30 > $O/lines_${W}_${KN}_$TAG.txt 2>&1
    done
  fi
done
tail -3 $O/pytest_gpu_$TAG.log; cut -c1-300 $O/bench_c2_$TAG.json; grep pls_gram $O/prof_C3_$TAG.txt; du -sh $O
