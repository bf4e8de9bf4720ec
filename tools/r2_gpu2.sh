#!/bin/bash
# render_tc bring-up: parity tests of the renderer (timeout-wrapped: a broken pipeline traps after 2^24 spins), then timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render" 2>&1 | tail -25 > gpurun_out/r2_render_tests.log
cat gpurun_out/r2_render_tests.log
timeout 120 python tools/prof_render.py 8 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generator" 2>&1 | tail -8
