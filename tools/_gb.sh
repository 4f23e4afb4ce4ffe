timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err
SCHEMANET_TABLE_ROWS=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r4.json 2>gpurun_out/bench.err
SCHEMANET_ADJ_TILED=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tiled.json 2>gpurun_out/bench.err
