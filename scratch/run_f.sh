set -x
timeout 900 python -m pytest tests -m gpu -q -x -k "head or umma or gemm or mn_major" 2>&1 | tail -3
timeout 300 python scratch/exp4.py 2>&1 | grep -v Warn | head -16
