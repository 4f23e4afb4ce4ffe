#!/bin/bash
# same-box A/B of an environment switch: AB_VAR=NAME (A: NAME=1, B: unset)
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in a b; do
  if [ $v = a ]; then export $AB_VAR=1; else unset $AB_VAR; fi
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ab_$v.json 2>/dev/null
  python - $v $rep <<'PY'
import json, sys
d=json.load(open("gpurun_out/ab_%s.json" % sys.argv[1])); k=d["kernels"]
sel = {n: round(k[n]["ms_per_launch"]*1e3,1) for n in k} if sys.argv[2] == "1" else {}
print(sys.argv[1], "cfg2 step %.4f ms" % d["ms_per_step"], "parity %.2e" % d["parity"]["logits_max_rel"] if d.get("parity") else "", sel)
PY
done
done
for v in a b; do
  if [ $v = a ]; then export $AB_VAR=1; else unset $AB_VAR; fi
  for c in ${AB_CONFIGS:-cfg3}; do
  timeout 500 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ab_${c}_$v.json 2>/dev/null
  python - $v $c <<'PY'
import json, sys
d=json.load(open("gpurun_out/ab_%s_%s.json" % (sys.argv[2], sys.argv[1])))
print(sys.argv[1], sys.argv[2], "step %.4f ms" % d["ms_per_step"])
PY
  done
done
unset $AB_VAR
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
