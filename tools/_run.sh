timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "render or decoder or generator_weight" 2>&1 | tail -3
timeout 100 python tools/prof_render_bwd.py 6 2>&1 | tail -1
