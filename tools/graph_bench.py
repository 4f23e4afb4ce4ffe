#!/usr/bin/env python
"""Stage-2 microbenchmark: sh_dev_instance_graphs alone at a named shape (CUDA events, L2 flushed between calls).
usage: tools/graph_bench.py [B] [M] [reps]   (env SCHEMANET_GRAPH_SPLIT / SCHEMANET_GRAPH_OCC select kernel variants)"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "schemanet-pytorch_b200"))
from schemanet_b200 import native  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
L = 196
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
ing = torch.randint(0, M, (B, L), device=dev, generator=g)
attn = 0.5 * torch.randn(B, L, L, device=dev, generator=g)
cls = 0.5 * torch.randn(B, L, device=dev, generator=g)
geo = torch.rand(L, L, device=dev, generator=g)
w = torch.full((2,), 0.5, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = native.PackedGraphs(B, L, dev, True, True)
ts = []
for i in range(reps + 3):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    native.instance_graphs(ing, attn, cls, geo, w, w, -1.0, -1.0, zero_pad=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
byts = B * (L * L * 4 * 2 + L * 12)   # attention read + padded edges written
print(f"B={B} M={M} split={os.environ.get('SCHEMANET_GRAPH_SPLIT', 'auto')} occ={os.environ.get('SCHEMANET_GRAPH_OCC', 'default')}: "
      f"median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us  ({byts / ts[len(ts) // 2] / 1e3:.0f} GB/s r+w)")
