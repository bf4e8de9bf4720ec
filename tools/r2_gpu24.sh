#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_b4.csv python tools/one_frame.py 2 --serial --batch 4 > gpurun_out/r2_ncu_list_b4.log 2>&1
tail -2 gpurun_out/r2_ncu_list_b4.log
timeout 600 python bench.py > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; tail -2 gpurun_out/r2_bench6.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench6.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['confirm'], d['sequential'], d['single_frame']); print(d['roofline']); print(d['render_roofline']['frac']); print(d['kernel_ms_per_frame'])
for k in ('train','train_rgb','reenact'): print(k, d[k]['value'], d[k].get('ms_per_step'), d[k].get('e2e'))"
