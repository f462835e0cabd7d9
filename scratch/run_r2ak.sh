#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 400 python scratch/host_floor.py 2>&1 | grep -E "^(1|8|16|24|32|64) " | cut -c1-260
