mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render or full_512 or generator_tiny" 2>&1 | tail -4
echo NEW; timeout 100 python tools/prof_render.py 8 2>&1 | tail -1
echo BASE; HFAGP_LIB=$PWD/hfa_gp_b200/libhfagp_base.so timeout 100 python tools/prof_render.py 8 2>&1 | tail -1
echo NEW; timeout 100 python tools/prof_render.py 8 2>&1 | tail -1
