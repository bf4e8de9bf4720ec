#!/bin/bash
for f in 2 4 8; do python bench.py --no-cpu-baseline --no-extras --frames-per-step $f 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('frames/step', $f, 'fps', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"; done
