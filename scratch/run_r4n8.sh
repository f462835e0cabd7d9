#!/bin/bash
# 8-GPU record at the end-of-round code: the driver's command
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_18_bench_n8.json 2> gpurun_out/r4n8.err
tail -c 300 gpurun_out/r4n8.err
