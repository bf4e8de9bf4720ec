#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/prof_upfir.py 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; tail -2 gpurun_out/r2_bench5.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['confirm'], d['sequential']); print(d['roofline']['frac'], d['render_roofline']['frac']); print(d['kernel_ms_per_frame'])
for k in ('train','train_rgb','reenact'): print(k, d[k]['value'], d[k].get('ms_per_step'), d[k].get('e2e'))"
timeout 600 python bench.py --workload train --tune-generator > gpurun_out/r2_bench5_tune.json 2> gpurun_out/r2_bench5_tune.err; tail -2 gpurun_out/r2_bench5_tune.err; cat gpurun_out/r2_bench5_tune.json | cut -c1-600
