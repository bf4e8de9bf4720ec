#!/bin/bash
# round-2 baseline: tests, bench, per-launch time + DRAM bytes, TF32 peak, ncu captures of the helper kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_pytest0.log
python bench.py --no-cpu-baseline > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
python tools/tf32_peak.py > gpurun_out/r2_tf32_peak.json 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches0.csv python tools/one_frame.py 2 --serial > gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:upfir_act -s 15 -c 1 -o gpurun_out/r2_upfir_sr1 python tools/one_frame.py 2 --serial >> gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blur_kernel -s 13 -c 1 -o gpurun_out/r2_blur python tools/one_frame.py 2 --serial >> gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 68 -c 5 -o gpurun_out/r2_conv_mid python tools/one_frame.py 2 --serial >> gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_epilogue -s 25 -c 1 -o gpurun_out/r2_conv_epi python tools/one_frame.py 2 --serial >> gpurun_out/r2_ncu_list.log 2>&1
ls -la gpurun_out | tail -12
cat gpurun_out/r2_pytest0.log gpurun_out/r2_bench0.json gpurun_out/r2_tf32_peak.json
