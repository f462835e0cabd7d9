#!/bin/bash
cd "$GRAFT_REPO_ROOT"
PNCE_EXPERIMENTS=1 timeout 300 python scratch/cfg4_timeline.py 8 2>&1 | tail -12
