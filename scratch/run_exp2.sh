set -x
timeout 600 python -m pytest tests/test_vs_eager_gpu.py -m gpu -q -s 2>&1 | tail -8
timeout 900 python scratch/config_sweep.py 2>&1 | grep -v Warn | tee gpurun_out/config_sweep.jsonl
