set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python scratch/exp15.py 2>&1 | grep -v Warn
python bench.py --no-cpu-baseline | cut -c1-200
