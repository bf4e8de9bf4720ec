#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "wgrad_tensor_core" 2>&1 | tail -15
