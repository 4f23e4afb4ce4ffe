#!/bin/bash
# GEMM checks: parity subset, per-role trace, cfg2 bench, cfg4 kernel table
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gnn or class_side or cfg2 or golden_fused" 2>&1 | tail -2
SCHEMANET_GEMM_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-graph 2>&1 >/dev/null | grep "gemm trace" | tail -12
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_c_cfg2.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_c_cfg2.json")); k=d["kernels"]
print("cfg2 step %.4f ms  %.0f img/s" % (d["ms_per_step"], d["value"]), {n: round(k[n]["ms_per_launch"]*1e3,1) for n in k if n.endswith("_tc")})
PY
if [ "$1" = "cfg4" ]; then
timeout 500 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_c_cfg4.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_c_cfg4.json")); k=d["kernels"]
print("cfg4 step %.3f ms  %.0f img/s" % (d["ms_per_step"], d["value"]))
for n,v in sorted(k.items(), key=lambda kv:-kv[1]["ms_total"])[:16]: print("  %-28s %3d x %8.1f us  %5.1f%%" % (n, v["launches"]//5, v["ms_per_launch"]*1e3, v["share"]*100))
PY
fi
