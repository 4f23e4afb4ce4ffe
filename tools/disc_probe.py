"""Stage-1 kernel times in isolation (profile API of the C library), cfg2 / cfg4 shapes.
Env: SCHEMANET_DISC_CTAS=2 (CTA pairs), SCHEMANET_DISC_DEBUG=1 (no epilogue) / 2 (no MMAs) bound the main kernel."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "schemanet-pytorch_b200"))
from schemanet_b200 import native

def run(R, d, M, reps=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(R, d, device="cuda", generator=g)
    c = torch.randn(M, d, device="cuda", generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        native.discretize(x, c)
    native.profile_enable(True)
    native.profile_collect()
    for _ in range(reps):
        flush.zero_()
        native.discretize(x, c)
    prof = native.profile_collect()
    native.profile_enable(False)
    tot = sum(ms for _, ms in prof.values()) / reps
    print(f"R={R} d={d} M={M}: stage {tot*1e3:.1f} us ", {k: round(ms / n * 1e3, 1) for k, (n, ms) in prof.items()},
          f" {2.0*R*d*M/tot/1e9:.0f} TFLOP/s whole stage")

if __name__ == "__main__":
    run(256 * 196, 384, 1024)
    if len(sys.argv) > 1:
        run(512 * 196, 768, 8000, reps=5)
