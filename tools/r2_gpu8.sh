#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -4
python bench.py --workload train --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rgb graph', d['ms_per_step'], d['value'], d['gpu_launches'])"
python bench.py --workload train --steps 20 --warmup 5 --no-graph 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rgb eager', d['ms_per_step'], d['value'], d['gpu_launches'])"
python bench.py --workload train --trainer 3dmm --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('3dmm b8 graph', d['ms_per_step'], d['value'], d['gpu_launches'])"
