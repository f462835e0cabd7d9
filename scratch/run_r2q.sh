#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nhwc_gpu.py tests/test_netf_gpu.py -m gpu -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2q_tests.log
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-head-line --layout nhwc"
timeout 300 $CMD > gpurun_out/r2q_plain_nhwc.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_loss_tc_p|k_gather_tc_nhwc|k_dense_nhwc|k_prep' -s 12 -c 4 -f -o gpurun_out/r2_02_nhwc $CMD > gpurun_out/r2q_ncu_nhwc.log 2>&1
echo "ncu nhwc rc=$?"; tail -2 gpurun_out/r2q_ncu_nhwc.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_02_nhwc_launches.csv $CMD > /dev/null 2>&1; echo "launch list rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-head-line --layout nhwc > gpurun_out/r2q_bench_nhwc.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_nhwc.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline','roofline_path','kernels_us'):
    print(k, json.dumps(d.get(k))[:700])
PY
