set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python scratch/exp4.py 2>&1 | grep -v Warn | head -12
timeout 300 python scratch/exp3.py 2>&1 | grep -v Warn | head -9
