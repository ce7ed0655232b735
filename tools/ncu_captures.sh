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
  set -- match_tc_kernel kmeans_persistent_kernel cond_phi_kernel kth_largest_kernel local_match_kernel affine_stats_partial \
         proxy_match_kernel upsample_softmax_label_kernel resize_bicubic2x_kernel
  # the unmasked (GCT statistics) and masked (FiLM pooling of the conditioning layer) passes are two instantiations of one
  # template: the first two launches of a frame are one of each (the details page names the instantiation)
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:channel_stats_partial -c 2 \
      -f -o $OUT/${P}_channel_stats_partial $CLIP > $OUT/${P}_channel_stats_partial_run.log 2>&1
  { echo "# ncu --set full --clock-control none -k regex:channel_stats_partial -c 2 $CLIP   (<(bool)0> = all pixels, <(bool)1> = phi > threshold)";
    grep -E "^(bank|frames)" $OUT/${P}_channel_stats_partial_run.log | sed 's/^/# /';
    ncu -i $OUT/${P}_channel_stats_partial.ncu-rep --page details; } > $OUT/${P}_channel_stats_partial_ncu_full.txt 2>&1
  ncu -i $OUT/${P}_channel_stats_partial.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw_pick.py >> $OUT/${P}_channel_stats_partial_ncu_full.txt
  # the convolution on its two largest layer types: halo variant (decoder conv1) and per-tap kernel (dilated decoder ASPP)
  for L in "dec.conv1" "dec.aspp"; do
    F=$(echo "$L" | tr -c 'A-Za-z0-9_\n' '_')
    ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 3 -c 1 -f -o $OUT/${P}_conv2_$F python tools/one_conv.py "$L" > /dev/null 2>&1
    { echo "# ncu --set full --clock-control none -k regex:conv2_kernel -s 3 -c 1 python tools/one_conv.py $L"; ncu -i $OUT/${P}_conv2_$F.ncu-rep --page details; } > $OUT/${P}_conv2_${F}_ncu_full.txt 2>&1
    ncu -i $OUT/${P}_conv2_$F.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw_pick.py >> $OUT/${P}_conv2_${F}_ncu_full.txt
  done
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
