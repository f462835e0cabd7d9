set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python bench.py --head --no-e2e --no-cpu-baseline > gpurun_out/bench_head.json 2>/dev/null
cat gpurun_out/bench_default.json; cut -c1-260 gpurun_out/bench_head.json
