#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --graph --batch 4 > gpurun_out/r2_timeline_graph_b4.txt 2>&1; grep -E "graph frame|last frame" gpurun_out/r2_timeline_graph_b4.txt; tail -24 gpurun_out/r2_timeline_graph_b4.txt
