#!/usr/bin/env python
import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.json"))
print({k: d[k] for k in ["value", "ms_per_step", "gpu_launches"]}, "e2e", round(d["e2e"]["value"]), "cached", round(d.get("class_side_cached", {}).get("value", 0)), d["clocks"])
r = d["roofline"]; print("roofline", r["kernel"], r["bound"], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3))
print("cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 1))
tot = sum(v["ms_total"] for v in d["kernels"].values()) / d["steps"]
for k, v in d["kernels"].items():
    print(f"{k:28s} {v['launches'] / d['steps']:4.1f}/step {v['ms_per_launch']:.4f} ms  share {v['share']:.3f}")
print("sum kernels/step ms", round(tot, 3)); print(d["stages"])
