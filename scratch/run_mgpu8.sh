set -x
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 300 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 --head --no-e2e > gpurun_out/bench_n${N}_head.json 2> gpurun_out/bench_n${N}_head.err; tail -c 300 gpurun_out/bench_n${N}_head.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_n${N}_ref.json 2>/dev/null
cut -c1-400 gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_head.json gpurun_out/bench_n${N}_ref.json
