#!/bin/bash
cd "$GRAFT_REPO_ROOT"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -o /tmp/compbench scratch/compbench.cu -lcuda 2>&1 | tail -5
timeout 120 /tmp/compbench 2>&1 | tail -8
