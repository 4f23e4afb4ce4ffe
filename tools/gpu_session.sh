#!/bin/bash
# One GPU session: TC discretize tests first (bounded), then the rest of the suite, bench, ncu captures.
mkdir -p gpurun_out
echo "== discretize tests (tensor-core path)" > gpurun_out/session.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "discretize or gnn" >> gpurun_out/session.log 2>&1
echo "rc=$?" >> gpurun_out/session.log
echo "== full gpu suite" >> gpurun_out/session.log
timeout 600 python -m pytest tests -m gpu -q >> gpurun_out/session.log 2>&1
echo "rc=$?" >> gpurun_out/session.log
echo "== bench" >> gpurun_out/session.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2>> gpurun_out/session.log
echo "rc=$?" >> gpurun_out/session.log
echo "== ncu launch list" >> gpurun_out/session.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/session.log
echo "== ncu full: graph build / discretize_tc / atlas" >> gpurun_out/session.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm3x|discretize_tc|instance_graph|class_edges|class_adj_raw" -s 9 -c 9 -o gpurun_out/prof_head python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/session.log
tail -60 gpurun_out/session.log
