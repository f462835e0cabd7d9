#!/bin/bash
# 2-GPU box at the end-of-round code (programmatic dependent launch on): every GPU test incl. the NCCL ones, bench N=2
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_15_gpu_tests_2gpu_box.log 2>&1; tail -n 2 gpurun_out/r2_15_gpu_tests_2gpu_box.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_15_bench_n2.json 2> gpurun_out/r4f_bench_n2.err
tail -c 300 gpurun_out/r4f_bench_n2.err
