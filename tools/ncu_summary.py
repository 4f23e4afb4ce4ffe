#!/usr/bin/env python
"""Summarise ncu outputs (read offline) into the text files kept under profiles/.

  tools/ncu_summary.py launches gpurun_out/launches.csv            -> per-kernel launch list / share of the step
  tools/ncu_summary.py full gpurun_out/prof.ncu-rep                -> key metrics of every captured launch
  tools/ncu_summary.py traffic gpurun_out/prof.ncu-rep             -> JSON {kernel: dram bytes / us per launch} (bench.py reads
                                                                     the newest profiles/rNN_ncu_traffic.json for roofline.traffic)
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "sm__cycles_elapsed.max"]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name, val = r[4], float(r[-1])
        short = name.split("(")[0].replace("void ", "").replace("sh::", "")
        if short.startswith("at::") or "at::native" in name:
            short = "[torch] " + short[:60]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches, {tot / 1e3:.1f} us total device time (cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'kernel':60s} {'launches':>8s} {'us/launch':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {n:8d} {t / n / 1e3:10.2f} {t / tot * 100:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[ik][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"    {k:75s} {r[i]:>18s} {units[i]}")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        print(f"    {'traffic = dram read + write (units as above)':75s} {float(r[rd]):.6g} {units[rd]} + {float(r[wr]):.6g} {units[wr]}")
        print()


def traffic(path):
    """Per kernel (template name without arguments): the launch with the LARGEST dram traffic = the class-side launch of a
    family, and the mean over its captured launches."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik, it = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}      # newer ncu builds name the units with or without 'second'
    agg = {}
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("sh::", "").split("<")[0]
        b = float(r[rd]) * scale.get(units[rd], 1.0) + float(r[wr]) * scale.get(units[wr], 1.0)
        us = float(r[it]) * tscale.get(units[it].replace("second", "s").replace("usecond", "us"), 1.0)
        agg.setdefault(name, []).append((b, us, r[ik][:80]))
    res = {}
    for name, v in agg.items():
        top = max(v, key=lambda x: x[0])
        res[name] = {"dram_bytes_per_launch": top[0], "us_per_launch_ncu": top[1], "launch": top[2],
                     "launches_captured": len(v), "mean_dram_bytes": sum(x[0] for x in v) / len(v), "source": path.split("/")[-1]}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
