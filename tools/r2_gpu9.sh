#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity.py -m gpu -x -q -s -k "full_512 or trainer_rgb_steps_match or full_size" 2>&1 | grep -E "per-pixel|full (rgb|3dmm)|passed|failed|Error|assert" | head -40
