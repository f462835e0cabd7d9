set -x
timeout 120 ./scratch/gatherbench | grep -E "ld.cg  |bulk"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
timeout 120 python scratch/exp7.py 2>&1 | grep -v Warn | tail -14
