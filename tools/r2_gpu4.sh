#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render" 2>&1 | tail -4
for v in 0 1; do echo "variant $v"; HFAGP_RT_VARIANT=$v timeout 120 python tools/prof_render.py 6 2>&1 | tail -1; done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:render_tc -s 2 -c 1 -o gpurun_out/r2_render_tc_v3 -f python tools/prof_render.py 4 > gpurun_out/r2_ncu_v3.log 2>&1
tail -4 gpurun_out/r2_ncu_v3.log
