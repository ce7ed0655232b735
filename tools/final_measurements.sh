#!/bin/bash
# Every measured artifact of a round in one GPU call (one B200): bench lines of all single-GPU configurations, layer tables,
# frame breakdown, launch list and ncu captures.  Outputs under gpurun_out/ with prefix $1 (default r2).
P=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 400 python bench.py > $O/${P}_bench_cfg3_480p_k5_n1.json 2> $O/${P}_bench_cfg3_n1.err
for c in cfg2 cfg4 cfg5; do timeout 300 python bench.py --config $c > $O/${P}_bench_$c.json 2> $O/${P}_bench_$c.err; done
timeout 200 python bench.py --impl reference --steps 4 > $O/${P}_bench_reference_arm.json 2>/dev/null
timeout 200 python tools/bench_conv.py > $O/${P}_bench_conv_layers.txt 2>&1
timeout 200 python tools/conv_table.py > $O/${P}_conv_table.txt 2>&1
timeout 100 python tools/frame_breakdown.py > $O/${P}_frame_breakdown.txt 2>&1
timeout 100 python tools/step_times.py > $O/${P}_step_times.txt 2>&1
AOCB200_LIB_TAG=trace timeout 100 python tools/conv_trace.py dec.conv1 dec.l1.conv3 > $O/${P}_conv_trace.txt 2>&1
AOCB200_LIB_TAG=trace timeout 100 python tools/conv_marks.py > $O/${P}_conv_marks.txt 2>&1
AOCB200_LIB_TAG=trace timeout 200 python tools/conv_attrib.py > $O/${P}_conv_attrib.txt 2>&1
timeout 900 tools/ncu_captures.sh $P
ls $O | wc -l
