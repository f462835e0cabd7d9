set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python scratch/exp6.py 2>&1 | grep -v Warn
timeout 600 python scratch/exp8.py 64 2>&1 | grep -v Warn
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
lscpu | head -20
