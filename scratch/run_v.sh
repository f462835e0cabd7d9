timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for s in 0 1 2 3; do timeout 250 python scratch/stress.py $s 200 2>&1 | grep -v Warn | tail -3; done
python bench.py --no-e2e --no-cpu-baseline --no-head-line | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench ms %.4f' % d['ms_per_step'], d['step_us'])"
