python bench.py --no-e2e --no-cpu-baseline --no-head-line | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fresh      ms %.4f' % d['ms_per_step'], d['kernels_us'])"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -1
python bench.py --no-e2e --no-cpu-baseline --no-head-line | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('after test ms %.4f' % d['ms_per_step'], d['kernels_us'])"
python bench.py --no-e2e --no-cpu-baseline --no-head-line | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('again      ms %.4f' % d['ms_per_step'], d['kernels_us'])"
nvidia-smi --query-gpu=temperature.gpu,temperature.memory,clocks.sm,clocks.mem,power.draw,memory.used --format=csv,noheader
