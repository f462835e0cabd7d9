set -x
N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 200 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 --head --no-e2e > gpurun_out/bench_n${N}_head.json 2> gpurun_out/bench_n${N}_head.err; tail -c 200 gpurun_out/bench_n${N}_head.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_n${N}_ref.json 2>/dev/null
for f in gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_head.json gpurun_out/bench_n${N}_ref.json; do python -c "
import json,sys
d=json.loads([l for l in open('$f') if l.startswith('{')][-1])
print('$f', d.get('impl','ours'), d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), d.get('step_us'), 'e2e', (d.get('e2e') or {}).get('value'))
"; done
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -q 2>&1 | tail -1
