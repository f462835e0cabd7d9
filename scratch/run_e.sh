set -x
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py tests/test_parity_gpu.py -q -x -k "two_gpu or head" 2>&1 | tail -4
N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 --head --no-e2e > gpurun_out/bench_n${N}_head.json 2> gpurun_out/bench_n${N}_head.err; tail -c 300 gpurun_out/bench_n${N}_head.err
cut -c1-330 gpurun_out/bench_n${N}_head.json
python bench.py --head --no-e2e --no-cpu-baseline --steps 100 | cut -c1-330
