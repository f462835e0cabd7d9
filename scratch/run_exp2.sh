set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5
python scratch/exp3.py 2>&1 | grep -v Warn
python bench.py --no-cpu-baseline --no-e2e
