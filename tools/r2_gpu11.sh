#!/bin/bash
for v in 1 0; do HFAGP_PDL=$v python bench.py --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('PDL', $v, 'fps', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'confirm', round(d['confirm']['value'],1))"; done
HFAGP_PDL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
