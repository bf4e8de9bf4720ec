#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -2 gpurun_out/r2_bench3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['confirm'], d['sequential']); print(d['roofline']['frac'], d['render_roofline']['frac'], d['cpu_baseline']);
for k in ('train','train_rgb','reenact'): print(k, d[k]['value'], d[k].get('ms_per_step'), d[k].get('e2e'))"
