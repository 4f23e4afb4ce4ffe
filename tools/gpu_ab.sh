#!/bin/bash
# same-box A/B of two builds of libschemahead.so (variant_a.so / variant_b.so, copied over the library in turn)
D=schemanet-pytorch_b200/schemanet_b200
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in ${AB_VARIANTS:-a b}; do
  cp $D/variant_$v.so $D/libschemahead.so
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ab_$v.json 2>/dev/null
  python - $v $rep <<'PY'
import json, sys
d=json.load(open("gpurun_out/ab_%s.json" % sys.argv[1])); k=d["kernels"]
sel = {n: round(k[n]["ms_per_launch"]*1e3,1) for n in k} if sys.argv[2] == "1" else {}
print(sys.argv[1], "cfg2 step %.4f ms" % d["ms_per_step"], sel)
PY
done
done
for v in ${AB_VARIANTS:-a b}; do
  cp $D/variant_$v.so $D/libschemahead.so
  for c in ${AB_CONFIGS:-cfg3 cfg4}; do
  timeout 500 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ab_${c}_$v.json 2>/dev/null
  python - $v $c <<'PY'
import json, sys
d=json.load(open("gpurun_out/ab_%s_%s.json" % (sys.argv[2], sys.argv[1])))
print(sys.argv[1], sys.argv[2], "step %.4f ms" % d["ms_per_step"])
PY
  done
done
cp $D/variant_${AB_FINAL:-b}.so $D/libschemahead.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gnn or class or cfg or golden or head or discretize" 2>&1 | tail -2
