#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 200 python scratch/graph_first_capture.py 1 320 > gpurun_out/r4r_graph_long.log 2>&1; grep -v "ACTIVE" gpurun_out/r4r_graph_long.log | tail -n 30
