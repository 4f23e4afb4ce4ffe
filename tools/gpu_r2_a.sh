#!/bin/bash
# Round-2 session A: parity first (new stage-1 band, full-size cfg2, training / init rows), then smoke and the bench line.
mkdir -p gpurun_out
L=gpurun_out/r2a.log
echo "== discretize tests" > $L
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "discretize" >> $L 2>&1
echo "rc=$?" >> $L
echo "== train / init tests" >> $L
timeout 400 python -m pytest tests/test_gpu_train.py -m gpu -q >> $L 2>&1
echo "rc=$?" >> $L
echo "== full gpu suite" >> $L
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_discretize_cta_pair_variant >> $L 2>&1
echo "rc=$?" >> $L
echo "== smoke" >> $L
timeout 120 python __graft_entry__.py smoke >> $L 2>&1
echo "rc=$?" >> $L
echo "== bench" >> $L
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2>> $L
echo "rc=$?" >> $L
tail -c 6000 $L
