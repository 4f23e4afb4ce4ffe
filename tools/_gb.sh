timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/disc_bench.py 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
