#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render or bookkeeping or generator_full" 2>&1 | tail -3
for v in 1 0; do echo "V2=$v"; HFAGP_RT_V2=$v timeout 120 python tools/prof_render.py 8 2>&1 | tail -1; done
