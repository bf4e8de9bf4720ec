#!/bin/bash
python tools/timeline.py --graph --seq 2>&1 | grep -v Warning | tail -175 > gpurun_out/r2_timeline_graph.txt; grep -E "graph frame|last frame" gpurun_out/r2_timeline_graph.txt
