#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload train --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rgb graph', d['ms_per_step'], d['value'])"
python tools/timeline_train.py 2>&1 | grep -E "one step|act_bwd|wgrad|render_kernel|conv_tc" 
