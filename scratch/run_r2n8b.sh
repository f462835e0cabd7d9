#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nproc; lscpu | grep -E "^CPU\(s\)|Thread|Core|Socket|Model name" 
for mode in pin nopin; do
  extra=""; [ $mode = nopin ] && extra="--no-pin"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e $extra > gpurun_out/r2_bench_n8_$mode.json 2> gpurun_out/r2_bench_n8_$mode.err; echo "bench $mode rc=$?"
  python - $mode <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r2_bench_n8_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'main', round(d['ms_per_step'],4), d['step_us'], 'cores/rank', d.get('host_cores_per_rank'))
print('  head', d['head_mode']['ms_per_step'], 'strong', d['strong']['ms_per_step'], d['strong']['direct']['ms_per_step'], 'nhwc', d['nhwc']['ms_per_step'], d['nhwc']['strong'])
PY
done
