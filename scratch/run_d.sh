for cp in 0.02 0 0.02 0; do
python bench.py --no-e2e --no-cpu-baseline --no-head-line --clock-period $cp | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('period $cp', d['ms_per_step'], d['clocks'])"
done
python bench.py --no-e2e --no-cpu-baseline --no-head-line --steps 1000 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1000 steps', d['ms_per_step'], d['clocks'])"
python scratch/exp9.py 64 2>&1 | grep persist=0 | head -1
