#!/bin/bash
python bench.py --workload train --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rgb', d['ms_per_step'], d['value'], d['gpu_launches'])"
python bench.py --workload train --trainer 3dmm --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('3dmm b8', d['ms_per_step'], d['value'], d['gpu_launches'])"
python bench.py --workload train --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rgb 5/3', d['ms_per_step'], d['value'], d['gpu_launches'])"
python tools/timeline_train.py 2>&1 | tail -45
