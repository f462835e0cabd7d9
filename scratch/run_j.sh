for i in 1 2 3 4 5 6; do
python bench.py --no-e2e --no-cpu-baseline --no-head-line | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('run $i ms_per_step %.4f dense_ms %.4f fwd+rest %.4f' % (d['ms_per_step'], d['roofline']['launch_ms'], d['ms_per_step']-d['roofline']['launch_ms']))"
done
