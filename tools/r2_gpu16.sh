#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -2
python bench.py --workload train --steps 10 --warmup 5 --tune-generator 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rgb tune', d['ms_per_step'], d['value'])"
python tools/timeline_train.py --tune 2>&1 | grep -E "one step|n= " | head -24
