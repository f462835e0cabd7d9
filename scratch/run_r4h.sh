#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_multicopy_gpu.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -n 40
timeout 200 python scratch/graph_first_capture.py 1 > gpurun_out/r4g_graph1.log 2>&1; tail -n 40 gpurun_out/r4g_graph1.log
