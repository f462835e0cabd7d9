set -x
timeout 300 python scratch/exp4.py 2>&1 | grep -v Warn | head -14
