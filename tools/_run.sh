mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/timeline_train_graph.py --trainer rgb --batch 2 --seq > gpurun_out/tl_rgb_b2.txt 2>&1; grep -v "^ *[0-9.]* " gpurun_out/tl_rgb_b2.txt | grep "busy\|pack_conv\|unpack_conv\|at::native\|split4"
timeout 300 python bench.py --workload train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train rgb', d['value'], d['ms_per_step'])"
