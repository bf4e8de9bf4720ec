#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline_train_ops.py > gpurun_out/r2_train_ops.txt 2>&1; tail -45 gpurun_out/r2_train_ops.txt
timeout 300 python tools/timeline_train.py 2>&1 | grep -E "one step|act_bwd"
