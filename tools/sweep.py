#!/usr/bin/env python
"""Microbenchmark sweep of BASELINE.json configs[4]: vocab 256-8192 x batch 1-4096 for the three head stages
(discretize, instance-graph build, instance-side match with the class side cached), d=384, K=100, D=256.
Per-stage times are the library's own CUDA-event kernel timings (sh_profile_enable), median of `--iters` steps.
usage: python tools/sweep.py [--out profiles/r01_sweep.md] [--quick]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "schemanet-pytorch_b200"), os.path.join(ROOT, "oracle"), ROOT):
    sys.path.insert(0, p)
import torch
import head_oracle as ho
from schemanet_b200 import native
from bench import build_head

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_sweep.md"))
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()

Ms = [256, 1024, 8192] if args.quick else [256, 512, 1024, 2048, 4096, 8192]
Bs = [1, 64, 1024] if args.quick else [1, 4, 16, 64, 256, 1024, 4096]
d, K, Vc, D, L = 384, 100, 256, 256, 196
dev = torch.device("cuda")
STAGES = {"discretize": ("rows_to_half_kernel", "codebook_norms_kernel", "row_sqnorm_kernel", "discretize_tc_kernel",
                         "discretize_tc_f16_kernel", "discretize_exact_kernel", "discretize_recheck_kernel"),
          "graph": ("instance_graph_kernel",),
          "match": ("gnn_adj_prep", "gnn_embed_gather", "gnn_embed_table_linear", "gnn_split_weights", "gnn_adj_gemm_tc",
                    "gnn_adj_ln_tc", "gnn_linear_ln_tc", "gnn_linear_tc", "gnn_ln_relu_wide", "gnn_pool_rows", "gnn_pool_fc",
                    "similarity_kernel", "gnn_adj_gemm", "gnn_linear_gemm", "ln_relu_kernel")}
lines = ["# Stage sweep on B200 (d=384, K=100, D=256; class side cached; images/s per stage)", "",
         "| M | B | discretize us | graph build us | instance match us | discretize Mimg/s | graph Mimg/s | match Mimg/s |",
         "|---|---|---|---|---|---|---|---|"]
for M in Ms:
    c = dict(B=1, d=d, H=6, M=M, K=K, Vc=min(Vc, M), D=D)
    schema = ho.synth_schema(M, K, c["Vc"], 5)
    gnn = ho.synth_gnn(M, D, 6)
    g = torch.Generator(device="cuda").manual_seed(M)
    vocab = torch.rand(M, d, device=dev, generator=g)
    head = build_head(c, vocab.cpu(), schema, gnn, dev)
    head.overlap_class_side = False
    for B in Bs:
        if B * M > 4096 * 8192 // 2 and args.quick:
            continue
        pick = torch.randint(0, M, (L + 1, B), device=dev, generator=g)
        mid = (vocab[pick] + 0.3 * torch.randn(L + 1, B, d, device=dev, generator=g)).contiguous()
        attn = 0.5 * torch.randn(B, L, L, device=dev, generator=g)
        cls = 0.5 * torch.randn(B, L, device=dev, generator=g)
        head(mid, attn, cls, cache_class=True)       # warm-up + class cache
        head(mid, attn, cls, cache_class=True)
        native.profile_enable(True)
        for _ in range(args.iters):
            head(mid, attn, cls, cache_class=True)
        prof = native.profile_collect()
        native.profile_enable(False)
        t = {s: sum(prof[k][1] for k in ks if k in prof) / args.iters * 1e3 for s, ks in STAGES.items()}
        lines.append(f"| {M} | {B} | {t['discretize']:.1f} | {t['graph']:.1f} | {t['match']:.1f} | "
                     f"{B / t['discretize']:.3f} | {B / t['graph']:.3f} | {B / t['match']:.3f} |")
        print(lines[-1], flush=True)
        del mid, attn, cls
    head._class_cache = None
open(args.out, "w").write("\n".join(lines) + "\n")
print("wrote", args.out)
