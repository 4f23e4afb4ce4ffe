#!/bin/bash
# Round-2 session B: GEMM with in-shared-memory operand split -- GNN tests first (bounded), then the suite and the bench.
mkdir -p gpurun_out
L=gpurun_out/r2b.log
echo "== gnn tests" > $L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gnn or class_side" >> $L 2>&1
echo "rc=$?" >> $L
echo "== full gpu suite" >> $L
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_discretize_cta_pair_variant >> $L 2>&1
echo "rc=$?" >> $L
echo "== bench" >> $L
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2>> $L
echo "rc=$?" >> $L
tail -c 5000 $L
