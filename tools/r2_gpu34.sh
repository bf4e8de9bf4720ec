#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blur_tile -s 1 -c 1 -o gpurun_out/r2_blur_b4 -f python tools/one_frame.py 1 --serial --batch 4 > gpurun_out/r2_ncu_blur.log 2>&1
tail -3 gpurun_out/r2_ncu_blur.log
