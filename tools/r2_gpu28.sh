#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
tail -3 gpurun_out/r2_bench_n8_final.err
python - <<'PY'
import json
s=open('gpurun_out/r2_bench_n8_final.json').read(); i=s.index('{"metric'); d=json.loads(s[i:].splitlines()[0])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'seq',d.get('sequential'),'single',d.get('single_frame'))
for k in ('train','train_rgb','reenact'):
    t=d[k]; print(k, t['value'], t.get('ms_per_step'), t.get('per_rank_batch'), t.get('allreduce_floats_per_step'))
PY
