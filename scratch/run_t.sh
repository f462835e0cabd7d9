timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for i in 1 2 3 4 5; do
python bench.py --no-e2e --no-cpu-baseline --no-head-line | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('run $i ms %.4f' % d['ms_per_step'], d['step_us'])"
done
