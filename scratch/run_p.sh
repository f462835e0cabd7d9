set -x
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -q -x 2>&1 | tail -4
