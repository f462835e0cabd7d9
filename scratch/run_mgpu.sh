set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 400 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --head --no-e2e --no-cpu-baseline > gpurun_out/bench_n2_head.json 2> gpurun_out/bench_n2_head.err
tail -c 400 gpurun_out/bench_n2_head.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_n2_ref.json 2>/dev/null
cat gpurun_out/bench_n2.json gpurun_out/bench_n2_head.json gpurun_out/bench_n2_ref.json | cut -c1-330
