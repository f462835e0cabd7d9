#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nhwc_gpu.py tests/test_netf_gpu.py -m gpu -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2r_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-head-line --layout nhwc > gpurun_out/r2r_bench_nhwc.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_bench_nhwc.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline_path','kernels_us'):
    print(k, json.dumps(d.get(k))[:700])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --head"
timeout 300 $CMD > gpurun_out/r2r_plain_head.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_loss_tc_p|k_gemm_tc_p|k_wgrad_tc' -s 18 -c 6 -f -o gpurun_out/r2_03_head $CMD > gpurun_out/r2r_ncu_head.log 2>&1
echo "ncu head rc=$?"; tail -2 gpurun_out/r2r_ncu_head.log
