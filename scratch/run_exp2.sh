set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --no-cpu-baseline --no-e2e | cut -c1-200
