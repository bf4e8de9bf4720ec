#!/bin/bash
# Profiling recipe of this repo on a B200 box (run under gpurun; one GPU): launch list with DRAM bytes at the bench's batch,
# warm in-graph timeline, ncu --set full of the top kernels, in-kernel role counters of the convolution.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_b4.csv python tools/one_frame.py 2 --serial --batch 4 > gpurun_out/ncu_list_b4.log 2>&1
timeout 300 python tools/timeline.py --graph --batch 4 > gpurun_out/timeline_graph_b4.txt 2>&1
# the captured training steps of bench.py's `train` (one frame per rank = the 8-GPU shard) and `train_rgb` records, node by node
timeout 300 python tools/timeline_train_graph.py --trainer 3dmm --batch 1 --seq > gpurun_out/timeline_train_graph_3dmm_b1.txt 2>&1
timeout 300 python tools/timeline_train_graph.py --trainer rgb --batch 2 --seq > gpurun_out/timeline_train_graph_rgb_b2.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_tc -s 2 -c 1 -o gpurun_out/render_tc -f python tools/prof_render.py 4 > gpurun_out/ncu_render.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o gpurun_out/conv_up_sr1 -f python tools/prof_conv.py --only sr1.conv0 --reps 3 > gpurun_out/ncu_conv_up.log 2>&1
timeout 120 python tools/prof_conv.py 2>&1 | tail -10
timeout 120 python tools/prof_upfir.py 2>&1 | tail -3
timeout 120 python tools/prof_blur.py 4 2>&1 | tail -13
# needs hfa_gp_b200/libhfagp_sm100_dbg.so (all csrc/*.cu compiled with -DHFAGP_TC_TIMING, linked like the Makefile does)
[ -f hfa_gp_b200/libhfagp_sm100_dbg.so ] && timeout 120 python tools/tc_timing.py 256 256 128 up 2>&1 | tail -7
