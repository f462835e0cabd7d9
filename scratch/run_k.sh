set -x
timeout 300 python scratch/exp16.py 2>&1 | grep -v Warn | head -5
timeout 100 python scratch/exp7.py 2>&1 | grep -v Warn | tail -12
