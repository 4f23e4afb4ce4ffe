timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 0 48; do SCHEMANET_ADJ_CTAS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$c.json 2>gpurun_out/bench.err; done
