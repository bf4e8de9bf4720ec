#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -3 gpurun_out/r2_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['confirm']); print(json.dumps(d.get('train'))); print(json.dumps(d.get('reenact')))"
