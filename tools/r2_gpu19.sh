#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render or bookkeeping" 2>&1 | tail -3
for tw in 0 1 2 3 4; do echo "tw_log2=$tw"; HFAGP_RT_TILE_W_LOG2=$tw timeout 120 python tools/prof_render.py 6 2>&1 | tail -1; done
