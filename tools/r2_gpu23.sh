#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 3" "1 4" "2 2" "2 3" "4 2" "4 3" "8 1" "8 2"; do
  set -- $cfg
  timeout 300 python bench.py --frames-per-step $1 --in-flight $2 --no-extras --no-cpu-baseline > gpurun_out/sweep_$1_$2.json 2> gpurun_out/sweep_$1_$2.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/sweep_$1_$2.json')); print('fps=$1 inflight=$2', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'confirm', round(d['confirm']['value'],1))" 2>&1 | tail -1
done
timeout 300 python tools/timeline.py --graph > gpurun_out/r2_timeline_graph2.txt 2>&1; tail -30 gpurun_out/r2_timeline_graph2.txt
