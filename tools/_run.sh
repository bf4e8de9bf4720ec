mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2c_n1.json 2> gpurun_out/bench_r2c_n1.err; tail -2 gpurun_out/bench_r2c_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2c_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['confirm']['value'], d['sequential']['value'], d['single_frame']['value']); print(d['roofline']['frac'], d['roofline']['tensor_pipe_frac'], d['render_roofline']['frac'], d['gpu_launches'], d['clocks']); print(d['kernel_ms_per_frame'])
for k in ('train','train_rgb','reenact'): print(k, d[k]['value'], d[k].get('ms_per_step'), d[k].get('gpu_launches_per_step'))"
timeout 300 python bench.py --workload train --tune-generator 2>/dev/null > gpurun_out/bench_r2c_tune.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2c_tune.json')); print('tune', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --workload train --trainer 3dmm --frames-per-step 1 2>/dev/null > gpurun_out/bench_r2c_3dmm_b1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2c_3dmm_b1.json')); print('3dmm b1', d['value'], d['ms_per_step'])"
cp gpurun_out/tl_rgb_b2.txt gpurun_out/tl_rgb_b2_prev.txt 2>/dev/null
timeout 300 python tools/timeline_train_graph.py --trainer 3dmm --batch 1 > gpurun_out/tl_3dmm_b1_final.txt 2>&1
timeout 300 python tools/timeline_train_graph.py --trainer rgb --batch 2 > gpurun_out/tl_rgb_b2_final.txt 2>&1
