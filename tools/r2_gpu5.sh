#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['render_roofline']['ms_per_launch']); print(d['kernel_ms_per_frame'])"
