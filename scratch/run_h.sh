set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 200 python scratch/exp12.py 1 2>&1 | grep -v Warn | head -3
timeout 200 python scratch/exp12.py 16 2>&1 | grep -v Warn | head -3
timeout 600 python scratch/config_sweep.py 2>&1 | grep -v Warn > gpurun_out/config_sweep.jsonl; cat gpurun_out/config_sweep.jsonl | cut -c1-260
