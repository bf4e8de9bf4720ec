#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render or bookkeeping or generator_full" 2>&1 | tail -3
timeout 120 python tools/prof_render.py 8 2>&1 | tail -1
