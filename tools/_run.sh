mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > gpurun_out/bench_r2c_n8.json 2> gpurun_out/bench_r2c_n8.err; tail -3 gpurun_out/bench_r2c_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r2c_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])
for k in ('train','train_rgb','reenact'): print(k, d[k]['value'], d[k].get('ms_per_step'), d[k].get('allreduce_floats_per_step'), d[k].get('final_loss'))"
