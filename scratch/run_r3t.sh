#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scratch/comp_ranks.py 2>&1 | grep "^rank"
timeout 300 python scratch/comp_ranks.py 2>&1 | grep "^rank"
