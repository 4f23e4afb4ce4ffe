#!/bin/bash
# compute-sanitizer over the round-2 kernels: memcheck on a parity subset (fp16 stage 1, fused class side, wide path, in-GEMM gather,
# tensor-core logits, training step), racecheck + synccheck on smoke and two small GEMM cases
mkdir -p gpurun_out
L=gpurun_out/r2_sanitizer.txt
echo "== memcheck: discretize (f16/tf32, outliers, edge cases), class side fused, gnn tensor-core paths (D=256 / wide), similarity, train step" > $L
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -x -q \
  -k "discretize_edge_cases or (discretize_outlier and f16 and 384) or class_side_fused or gnn_tensor_core_path or gnn_wide_fused or (similarity_tensor and 257) or train_step or golden_fused" >> $L 2>&1
echo "memcheck rc=$?" >> $L
echo "== racecheck: smoke" >> $L
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke >> $L 2>&1
echo "racecheck smoke rc=$?" >> $L
echo "== racecheck: wide fused path (in-GEMM gather, LayerNorm in the operand conversion), fused class side" >> $L
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(gnn_wide_fused and 100) or (class_side_fused and 296)" >> $L 2>&1
echo "racecheck tests rc=$?" >> $L
echo "== synccheck: smoke" >> $L
timeout 200 compute-sanitizer --tool synccheck --error-exitcode 9 python __graft_entry__.py smoke >> $L 2>&1
echo "synccheck smoke rc=$?" >> $L
grep -v "^$" $L | tail -40
