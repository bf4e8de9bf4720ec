#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/prof_blur.py 4 2>&1 | tail -13
timeout 300 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sequential']['value'], d['single_frame']); print(d['kernel_ms_per_frame'])"
