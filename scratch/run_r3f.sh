#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scratch/split_breakdown.py 64 2>&1 | grep -v Warn | cut -c1-900
timeout 300 python scratch/split_breakdown.py 16 2>&1 | grep -v Warn | cut -c1-900
