for k in 1 2 3 4 5 6 7 8; do timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q -k "audio" 2>&1 | grep -v "^E   *where\|^E   *+" | tail -4 | grep -v "^$"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "driven or frame_loop" 2>&1 | tail -3
