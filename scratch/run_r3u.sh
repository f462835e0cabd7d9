#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for c in 1 0; do echo "PNCE_GRAD_COMPRESSION=$c"; PNCE_GRAD_COMPRESSION=$c timeout 300 python scratch/cfg4_breakdown.py 2>&1 | grep -v Warn | grep -E "^(nchw|nhwc)" | cut -c1-260; done
