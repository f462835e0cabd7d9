set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 600 python scratch/config_sweep.py 2>&1 | grep -E "cfg4|cfg1/2: B=1, 256\^2, B5|cfg5 per-GPU: B=64|cfg3: B=16, 256\^2, B5" | cut -c1-260
