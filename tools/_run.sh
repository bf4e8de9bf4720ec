mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_tc -s 2 -c 1 -o gpurun_out/render_tc_v4 -f python tools/prof_render.py 4 > gpurun_out/ncu_render.log 2>&1; tail -2 gpurun_out/ncu_render.log
ls -la gpurun_out/*.ncu-rep
