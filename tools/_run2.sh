timeout 300 python tools/timeline_train_ops.py --3dmm --batch 1 2>&1 | tail -45
