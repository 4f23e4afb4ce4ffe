#!/usr/bin/env python
"""Top SASS instructions of a kernel by warp-stall samples, from an .ncu-rep (read offline, no GPU needed).
usage: tools/ncu_hot.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + regex],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for n, r in enumerate(rows[2:]):
    if len(r) <= isamp or r[0] == "Kernel Name":
        break
    try:
        data.append((n, float(r[isamp]), float(r[iexec]), r[isrc].strip()))
    except ValueError:
        pass
tot = sum(d[1] for d in data) or 1
print(f"{len(data)} instructions, {tot:.0f} samples, {sum(d[2] for d in data):.3g} warp-instructions executed")
for n, s, e, src in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{s / tot * 100:5.1f}%  #{n:5d} exec={e:10.0f}  {src[:100]}")
