#!/bin/bash
# TMA gather4 producer of the wide path's layer-0 GEMM: parity, then cfg4 A/B against the per-node bulk copies (SCHEMANET_GEMM_DEBUG=8)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or cfg4 or more_tiles or golden" 2>&1 | tail -2
for v in bulk gather4 bulk gather4; do
  if [ $v = bulk ]; then export SCHEMANET_GEMM_DEBUG=8; else unset SCHEMANET_GEMM_DEBUG; fi
  timeout 500 python bench.py --config cfg4 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/g4_$v.json 2>/dev/null
  python - $v <<'PY'
import json, sys
d=json.load(open("gpurun_out/g4_%s.json" % sys.argv[1])); k=d["kernels"]
print(sys.argv[1], "cfg4 step %.3f ms  %.0f img/s" % (d["ms_per_step"], d["value"]), {n: round(k[n]["ms_per_launch"]*1e3,1) for n in k if n.endswith("_tc")})
PY
done
