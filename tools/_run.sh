mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q -k "basis_qr or latent or full_size or 3dmm_step" 2>&1 | tail -3
timeout 300 python tools/timeline_train_graph.py --trainer 3dmm --batch 1 --seq > gpurun_out/tl_3dmm_b1.txt 2>&1; grep "busy\|batch 1:" gpurun_out/tl_3dmm_b1.txt; grep " qr_" gpurun_out/tl_3dmm_b1.txt | head -8
