#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scratch/head_small_breakdown.py 1 8 16 64 2>&1 | grep -v Warn | cut -c1-600
