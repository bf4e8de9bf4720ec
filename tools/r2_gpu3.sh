#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_tc -s 2 -c 1 -o gpurun_out/r2_render_tc_v3 -f python tools/prof_render.py 4 > gpurun_out/r2_ncu_v3.log 2>&1
tail -15 gpurun_out/r2_ncu_v3.log
