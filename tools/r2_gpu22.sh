#!/bin/bash
mkdir -p gpurun_out
echo "== blur auto"; timeout 120 python tools/prof_blur.py 2>&1 | tail -13
echo "== blur big=0"; HFAGP_BLUR_BIG=0 timeout 120 python tools/prof_blur.py 2>&1 | tail -13
echo "== blur big=1"; HFAGP_BLUR_BIG=1 timeout 120 python tools/prof_blur.py 2>&1 | tail -13
echo "== conv"; timeout 200 python tools/prof_conv.py 2>&1 | tail -10
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o gpurun_out/r2_conv_up_sr1 -f python tools/prof_conv.py --only sr1.conv0 --reps 3 > gpurun_out/r2_ncu_up.log 2>&1
tail -3 gpurun_out/r2_ncu_up.log
