#!/bin/bash
for f in 1 2 3; do python bench.py --no-cpu-baseline --no-extras --in-flight $f 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('in-flight', $f, 'fps', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'confirm', round(d['confirm']['value'],1))"; done
