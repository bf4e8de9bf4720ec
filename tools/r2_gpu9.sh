#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity.py -m gpu -x -q -k "full_size or bookkeeping" 2>&1 | tail -30
