for i in 1 2 3; do
for st in 0 6; do
python bench.py --no-e2e --no-cpu-baseline --no-head-line --settle $st | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('run $i settle $st ms_per_step %.4f dense_ms %.4f fwd+rest %.4f' % (d['ms_per_step'], d['roofline']['launch_ms'], d['ms_per_step']-d['roofline']['launch_ms']))"
done
done
python bench.py --no-e2e --no-cpu-baseline --no-head-line --warmup 400 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warmup 400 ms_per_step %.4f dense_ms %.4f' % (d['ms_per_step'], d['roofline']['launch_ms']))"
python bench.py --no-e2e --no-cpu-baseline --no-head-line --steps 2000 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('steps 2000 ms_per_step %.4f dense_ms %.4f' % (d['ms_per_step'], d['roofline']['launch_ms']))"
