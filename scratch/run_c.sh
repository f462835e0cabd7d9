set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python scratch/exp10.py 64 2>&1 | grep -v Warn
timeout 300 python scratch/exp9.py 64 2>&1 | grep -v Warn
timeout 300 python scratch/exp9.py 16 2>&1 | grep -v Warn
