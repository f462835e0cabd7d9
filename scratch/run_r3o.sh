#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scratch/exp51.py 2>&1 | tail -8
