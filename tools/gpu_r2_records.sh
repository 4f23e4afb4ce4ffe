#!/bin/bash
# Round-2 records: full GPU suite, smoke, bench (ours + reference arm), ncu launch list and full capture of the hot kernels
mkdir -p gpurun_out
L=gpurun_out/r2_records.log
echo "== full gpu suite" > $L
timeout 1200 python -m pytest tests -m gpu -q >> $L 2>&1
echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 120 python __graft_entry__.py smoke >> $L 2>&1
echo "rc=$?" >> $L
echo "== bench" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2>> $L
echo "rc=$?" >> $L
echo "== reference arm" >> $L
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> $L
echo "rc=$?" >> $L
echo "== ncu launch list" >> $L
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-graph > /dev/null 2>> $L
echo "== ncu full" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm3x|discretize_tc|instance_graph|class_edges|class_adj_raw|rows_to_half|embed_gather" -s 14 -c 14 -f -o gpurun_out/r2_hot python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-graph > /dev/null 2>> $L
tail -c 3000 $L
