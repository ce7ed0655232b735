#!/bin/bash
# Run on the GPU box (gpurun): launch list of one steady-state frame + one `ncu --set full` capture per kernel the
# north star names (B200_PROFILING.md recipe).  Outputs under gpurun_out/ with the prefix given as $1 (default r2).
#   tools/ncu_captures.sh [prefix] [kernel-regex ...]
P=${1:-r2}; shift
OUT=gpurun_out
mkdir -p $OUT
CLIP="python tools/run_clip.py --frames 1 --warm 7"
if [ $# -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file $OUT/${P}_launches_480p_k5.csv $CLIP > $OUT/${P}_launches_run.log 2>&1
  python tools/summarize_launches.py $OUT/${P}_launches_480p_k5.csv > $OUT/${P}_launches_480p_k5_summary.txt 2>&1
  set -- match_tc_kernel kmeans_persistent_kernel cond_phi_kernel "channel_stats_partial<true>" "channel_stats_partial<false>" \
         kth_largest_kernel local_match_kernel affine_stats_partial proxy_match_kernel upsample_softmax_label_kernel
fi
for K in "$@"; do
  F=$(echo "$K" | tr -c 'A-Za-z0-9_\n' '_')
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$K" -c 1 \
      -f -o $OUT/${P}_${F} $CLIP > $OUT/${P}_${F}_run.log 2>&1
  { echo "# ncu --set full --clock-control none -k regex:$K -c 1 $CLIP"; grep -E "^(bank|frames)" $OUT/${P}_${F}_run.log | sed 's/^/# /';
    ncu -i $OUT/${P}_${F}.ncu-rep --page details; } > $OUT/${P}_${F}_ncu_full.txt 2>&1
  # per-launch DRAM traffic (roofline.traffic of bench.py)
  ncu -i $OUT/${P}_${F}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw_pick.py >> $OUT/${P}_${F}_ncu_full.txt
done
