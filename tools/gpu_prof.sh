mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"instance_graph" -s 1 -c 1 -o gpurun_out/prof_graph python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu.err; tail -3 gpurun_out/ncu.err
