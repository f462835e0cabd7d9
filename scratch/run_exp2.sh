set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python scratch/exp4.py 2>&1 | grep -v Warn
MATH=tc_bf16 timeout 300 python scratch/exp4.py 2>&1 | grep -v Warn
