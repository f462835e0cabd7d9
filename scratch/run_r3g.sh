#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scratch/split_breakdown.py 16 prof > gpurun_out/r3g_split_prof.log 2>&1
grep -E "use_mlp|ncalls|patchnce.py|torch|built-in|method" gpurun_out/r3g_split_prof.log | cut -c1-170 | head -120
