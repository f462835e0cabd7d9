#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scratch/host_nhwc.py 2>&1 | grep -v "^$" | cut -c1-150 | head -70
