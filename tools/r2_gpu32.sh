#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -x -q -k "register_tiled or upconv or small_ops or decoder_weight_gradient or act_bwd or bookkeeping or render_matches_oracle" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -k "decoder_weight_gradient or act_bwd" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck.log
