mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --workload train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train rgb', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --workload train --trainer 3dmm --frames-per-step 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train 3dmm b1', d['value'], d['ms_per_step'])"
