#!/bin/bash
# End-of-round check on a B200 box (run under gpurun): GPU tests, smoke, the default bench line, the unfrozen training step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; tail -2 gpurun_out/bench_verify.err
python -c "
import json; d=json.load(open('gpurun_out/bench_verify.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['confirm']['value'], d['sequential']['value'], d['single_frame']['value']); print(d['roofline']['frac'], d['roofline']['tensor_pipe_frac'], d['render_roofline']['frac'], d['gpu_launches'], d['clocks']); print(d['kernel_ms_per_frame'])
for k in ('train','train_rgb','reenact'): print(k, d[k]['value'], d[k].get('ms_per_step'), d[k].get('e2e'))
print(d['cpu_baseline'])"
timeout 300 python bench.py --workload train --tune-generator 2>/dev/null > gpurun_out/bench_verify_tune.json; python -c "
import json; d=json.load(open('gpurun_out/bench_verify_tune.json')); print('tune', d['value'], d['ms_per_step'])"
