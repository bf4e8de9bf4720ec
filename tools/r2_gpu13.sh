#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1]); print('N8 fps', d['value'], 'e2e', d['e2e']['value'], 'confirm', d['confirm']); print(json.dumps(d.get('train'))); print(json.dumps(d.get('reenact')))"
