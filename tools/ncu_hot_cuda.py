#!/usr/bin/env python
"""Top CUDA-C source lines of a kernel by warp-stall samples (needs -lineinfo and --import-source on).
usage: tools/ncu_hot_cuda.py report.ncu-rep kernel_regex [top_n] [launch_index]"""
import csv
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv", "-k", "regex:" + regex],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches may follow each other: split at "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
for bi, b in enumerate(blocks):
    if bi != which:
        continue
    hdr = b["rows"][0]
    isrc = hdr.index("Source")
    isamp = hdr.index("# Samples")
    data = []
    for n, r in enumerate(b["rows"][1:]):
        try:
            data.append((float(r[isamp]), n, r[isrc].strip(), r[0]))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    print(f"[{bi}] {b['name'][:100]}: {tot:.0f} samples")
    for s, n, src, ln in sorted(data, key=lambda d: -d[0])[:top]:
        print(f"  {s / tot * 100:5.1f}%  L{ln:>5s}  {src[:120]}")
print(f"({len(blocks)} launches in report)")
