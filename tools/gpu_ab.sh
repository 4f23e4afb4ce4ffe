#!/bin/bash
# parity subset + cfg2 / cfg3 / cfg4 step times of the current build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -x -q -k "gnn or class or cfg or golden or head or train or similarity" 2>&1 | tail -2
for c in cfg2 cfg2 cfg3 cfg4; do
  timeout 500 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ab_$c.json 2>/dev/null
  python - $c <<'PY'
import json, sys
d=json.load(open("gpurun_out/ab_%s.json" % sys.argv[1])); k=d["kernels"]
print(sys.argv[1], "step %.4f ms  %.0f img/s" % (d["ms_per_step"], d["value"]), {n: round(k[n]["ms_per_launch"]*1e3,1) for n in k if n.endswith("_tc")})
PY
done
