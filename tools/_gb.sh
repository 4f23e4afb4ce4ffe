timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -c 600 gpurun_out/bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.json 2>gpurun_out/bench.err
