set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python scratch/exp18.py 2>&1 | grep -v Warn
