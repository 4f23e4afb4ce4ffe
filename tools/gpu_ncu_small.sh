#!/bin/bash
# ncu --set full (with source) of the small kernels of a step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_groups_fc|embed_gather|rows_linear|pool_table_rows|param_bounds|class_perm|adj_sym|similarity_kernel|class_vertices" -s 9 -c 14 -f -o gpurun_out/r2_small python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-graph > /dev/null 2> gpurun_out/r2_small.log
tail -3 gpurun_out/r2_small.log
