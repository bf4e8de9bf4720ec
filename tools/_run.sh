mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -2 gpurun_out/bench_r2b.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2b.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['confirm']['value'], d['sequential']['value'], d['single_frame']['value']); print(d['roofline']['frac'], d['roofline']['tensor_pipe_frac'], d['render_roofline']['frac'], d['gpu_launches'], d['clocks']); print(d['kernel_ms_per_frame'])
for k in ('train','train_rgb','reenact'): print(k, d[k]['value'], d[k].get('ms_per_step'), d[k].get('gpu_launches_per_step'))
print(d['cpu_baseline'])"
