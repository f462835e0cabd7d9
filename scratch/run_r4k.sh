#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
# is the slow-gather effect of a programmatically launched dense kernel tied to compressible gradient memory?
for c in 1 0; do for v in 23 31; do
  echo -n "PNCE_GRAD_COMPRESSION=$c PNCE_PDL=$v  "
  PNCE_GRAD_COMPRESSION=$c PNCE_PDL=$v timeout 200 python scratch/pdl_ab.py 2>&1 | tail -1
done; done > gpurun_out/r4k_pdl_compression.log 2>&1
cat gpurun_out/r4k_pdl_compression.log
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r4k_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4k_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['kernels_us']); print(d['roofline_other_kernels'][1]); print(d['nhwc']['kernels_us']); print(d['nhwc']['head_mode']['kernels_us'])
PY
