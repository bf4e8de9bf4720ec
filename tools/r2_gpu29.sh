#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "register_tiled" 2>&1 | tail -3
timeout 300 python tools/timeline_train.py > gpurun_out/r2_timeline_train2.txt 2>&1; tail -45 gpurun_out/r2_timeline_train2.txt
