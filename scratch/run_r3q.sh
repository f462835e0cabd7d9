#!/bin/bash
cd "$GRAFT_REPO_ROOT"
PNCE_EXPERIMENTS=1 timeout 300 python scratch/exp50.py 64 nchw 2>&1 | grep -v Warn | grep compress
