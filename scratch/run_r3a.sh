#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scratch/host_head.py 1 8 16 64 > gpurun_out/r3a_host_head.log 2>&1
tail -60 gpurun_out/r3a_host_head.log | cut -c1-200
