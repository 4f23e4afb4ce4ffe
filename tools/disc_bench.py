#!/usr/bin/env python
"""Times stage 1 alone (CUDA events, L2 flushed between iterations) for one or more shapes.
usage: tools/disc_bench.py [B d M mode]...   mode in {auto, exact, tf32, f16}"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schemanet-pytorch_b200"))
import torch
from schemanet_b200 import native

MODES = {"auto": native.DISC_AUTO, "exact": native.DISC_EXACT, "tf32": native.DISC_TENSOR, "f16": native.DISC_TENSOR_F16}
args = sys.argv[1:] or ["256", "384", "1024", "auto"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(0, len(args), 4):
    B, d, M, mode = int(args[i]), int(args[i + 1]), int(args[i + 2]), args[i + 3]
    g = torch.Generator(device="cuda").manual_seed(1)
    vocab = torch.rand(M, d, device="cuda", generator=g)
    x = vocab[torch.randint(0, M, (196 * B,), device="cuda", generator=g)] + 0.3 * torch.randn(196 * B, d, device="cuda", generator=g)
    out = torch.empty(196 * B, dtype=torch.int64, device="cuda")
    for _ in range(3):
        native.discretize(x, vocab, out_idx=out, mode=MODES[mode])
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); native.discretize(x, vocab, out_idx=out, mode=MODES[mode]); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    fl = 2.0 * 196 * B * d * M
    print(f"B={B} d={d} M={M} mode={mode} debug={os.environ.get('SCHEMANET_DISC_DEBUG','0')}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s  {196*B/ms*1e3/1e6:.2f} Mrows/s")
