#!/bin/bash
# ncu --set full of the GEMM launches of one bench step (class-side adjacency + LN, instance side, linear)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm3x" -s 12 -c 6 -f -o gpurun_out/r2_gemm \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-graph > /dev/null 2> gpurun_out/r2_ncu.err
tail -3 gpurun_out/r2_ncu.err
ls -la gpurun_out/r2_gemm.ncu-rep
