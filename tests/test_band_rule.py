"""CPU emulation of stage 1's candidate band (csrc/discretize_tc.cu) on non-i.i.d. features.

The tensor-core pass scores fp16-rounded operands; a codeword survives to the exact fp32 re-check when its coarse score is
within `band` of the coarse row minimum.  The band is a worst-case bound (Cauchy-Schwarz on the measured rounding
residuals + the fp32 accumulation error), so the exact winner must ALWAYS be short-listed -- in particular on features
with outlier channels (DeiT layer-9 "massive activations"), where a statistical band loses it.
This file restates the rule in numpy and checks that claim; the kernel itself is checked by
tests/test_gpu_parity.py::test_discretize_outlier_channels on the GPU.
"""
import numpy as np
import pytest


def band_rule(x, c, d):
    """The quantities the conversion pass measures and the band the epilogue derives from them (fp32 like the kernel)."""
    xh = np.clip(x, -65504, 65504).astype(np.float16).astype(np.float32)
    ch = np.clip(c, -65504, 65504).astype(np.float16).astype(np.float32)
    xn2 = (x.astype(np.float32) ** 2).sum(1, dtype=np.float32)
    xe2 = ((x - xh) ** 2).sum(1, dtype=np.float32) * np.float32(1.0000005)
    cn2 = (c.astype(np.float32) ** 2).sum(1, dtype=np.float32)
    ce2 = ((c - ch) ** 2).sum(1, dtype=np.float32) * np.float32(1.0000005)
    cmax2, dcmax2 = cn2.max(), ce2.max()
    cmax, dcmax = np.sqrt(cmax2), np.sqrt(dcmax2)
    nx, dx = np.sqrt(xn2), np.sqrt(xe2)
    nxh = nx + dx
    gamma = np.float32(2.0 * d * 2.0 ** -23)
    e = dx * cmax + nxh * dcmax + gamma * nxh * (cmax + dcmax)
    band = np.float32(4.0) * e + np.float32(2.0 ** -19) * (xn2 + cmax2)
    return xh, ch, cn2, band.astype(np.float32)


def outlier_case(d, M, R, scales, mode, seed):
    g = np.random.default_rng(seed)
    vocab = g.random((M, d), dtype=np.float32)
    if mode == "easy":
        x = vocab[g.integers(0, M, R)] + 0.3 * g.standard_normal((R, d), dtype=np.float32)
    else:
        x = g.random((R, d), dtype=np.float32)
    for ch, s in scales:                     # the same channels are large in the tokens AND in the codebook
        x[:, ch] *= s
        vocab[:, ch] *= s
    return x, vocab


CASES = [
    (384, 1024, 4000, [(7, 30.0)], "easy"),
    (384, 1024, 4000, [(7, 30.0)], "hard"),
    (384, 1024, 4000, [(7, 100.0), (200, 100.0)], "hard"),
    (768, 8000, 600, [(5, 30.0)], "hard"),
    (384, 1024, 2000, [], "hard"),
]


@pytest.mark.parametrize("d,M,R,scales,mode", CASES)
def test_exact_winner_is_always_short_listed(d, M, R, scales, mode):
    x, c = outlier_case(d, M, R, scales, mode, seed=d + M + len(scales))
    xh, ch, cn2, band = band_rule(x, c, d)
    # coarse pass: exact products of the rounded operands, fp32 accumulation (BLAS order stands in for the tensor core's)
    coarse = cn2[None, :] - np.float32(2.0) * (xh @ ch.T)
    # exact scores in fp64
    x64, c64 = x.astype(np.float64), c.astype(np.float64)
    exact = (c64 ** 2).sum(1)[None, :] - 2.0 * (x64 @ c64.T)
    win = exact.argmin(1)
    listed = coarse <= (coarse.min(1) + band)[:, None]
    lost = ~listed[np.arange(R), win]
    assert lost.sum() == 0, f"{lost.sum()} of {R} rows lose the exact winner"
    # and the theorem's inequality itself: |coarse - exact| <= 2e <= band / 2 on every entry
    err = np.abs(coarse.astype(np.float64) - exact)
    assert (err <= 0.5 * band[:, None].astype(np.float64)).all()
    per_row = listed.sum(1)
    print(f"d={d} M={M} {mode} scales={scales}: candidates/row mean {per_row.mean():.2f} max {per_row.max()}, "
          f"rows with > 1 candidate {100.0 * (per_row > 1).mean():.1f} %, rows with > 14 {100.0 * (per_row > 14).mean():.2f} %")


def test_bf16_statistical_band_does_lose_winners():
    """The round-1 rule (bf16 operands, band = 2^-8 |x| max|c|) on the same outlier data: documents why it was replaced."""
    import torch
    d, M, R = 384, 1024, 4000
    x, c = outlier_case(d, M, R, [(7, 30.0)], "hard", seed=3)
    xb = torch.from_numpy(x).bfloat16().float().numpy()
    cb = torch.from_numpy(c).bfloat16().float().numpy()
    cn2 = (c ** 2).sum(1, dtype=np.float32)
    coarse = cn2[None, :] - np.float32(2.0) * (xb @ cb.T)
    band = np.float32(2.0 ** -8) * np.sqrt((x ** 2).sum(1)) * np.sqrt(cn2.max())
    x64, c64 = x.astype(np.float64), c.astype(np.float64)
    exact = (c64 ** 2).sum(1)[None, :] - 2.0 * (x64 @ c64.T)
    srt = np.sort(exact, 1)
    unambiguous = (srt[:, 1] - srt[:, 0]) > 1e-6 * np.abs(srt[:, 0])
    listed = coarse <= (coarse.min(1) + band)[:, None]
    lost = ~listed[np.arange(R), exact.argmin(1)] & unambiguous
    assert lost.sum() > 0
