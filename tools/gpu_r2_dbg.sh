#!/bin/bash
# where does a GEMM k-block's time go?  bench with parts of gemm3x_kernel disabled (results are garbage: timing only)
mkdir -p gpurun_out
for d in 0 1 2 4 3 7; do
  SCHEMANET_GEMM_DEBUG=$d timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_dbg_$d.json 2> gpurun_out/r2_dbg_$d.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_dbg_$d.json'))
k=d['kernels']
print('debug=$d  step %.4f ms  ' % d['ms_per_step'] + '  '.join('%s %.1f' % (n, k[n]['ms_per_launch']*1e3) for n in ('gnn_adj_ln_tc','gnn_adj_gemm_tc','gnn_linear_ln_tc') if n in k))
PY
done
