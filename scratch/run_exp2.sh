set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15
python scratch/exp2.py 2>&1 | grep -v Warning
