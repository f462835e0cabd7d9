set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 100 python scratch/exp19.py 2>&1 | grep -v Warn
