#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/timeline_train.py 2>&1 | grep -E "one step|act_bwd|render_kernel"
timeout 300 python bench.py --workload train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train_rgb', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --workload train --tune-generator 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tune', d['value'], d['ms_per_step'])"
