"""Drop-in for `schema_inference.graph.utils` (schema_inference/graph/utils.py:8-106): normalisation helpers, the
geometric-similarity table of the 14x14 token grid and the `MyParameter` holder whose `.tensor` naming fixes the
state-dict keys (`vertex_weights.tensor`, ...).  Host-side glue: small tensors, stock torch ops."""
from typing import Iterable

import torch
import torch.nn as nn


def _div_nan0(x: torch.Tensor, denom: torch.Tensor, inplace: bool) -> torch.Tensor:
    if inplace:
        x /= denom
        return x.nan_to_num_(0)
    return (x / denom).nan_to_num(0)


def normalize_sum_(x: torch.Tensor, dim: int = -1):
    """In place: x / x.sum(dim), NaN -> 0."""
    return _div_nan0(x, x.sum(dim=dim, keepdim=True), True)


def normalize_max_(x: torch.Tensor, dim: int = -1):
    """In place: x / x.max(dim), NaN -> 0."""
    return _div_nan0(x, x.max(dim=dim, keepdim=True)[0], True)


def normalize_sum(x: torch.Tensor, dim: int = -1, detach_sum: bool = False):
    total = x.sum(dim=dim, keepdim=True)
    return _div_nan0(x, total.detach() if detach_sum else total, False)


def normalize_max(x: torch.Tensor, dim: int = -1):
    return _div_nan0(x, x.max(dim=dim, keepdim=True)[0], False)


def normalize_sum_clamp(x: torch.Tensor, dim: int = -1, detach_sum: bool = False, min_val: float = 0) -> torch.Tensor:
    return normalize_sum(x.clamp_min(min_val), dim, detach_sum=detach_sum)


def pair_wise_point_dist(h: int, w: int, pow: float = 2, device: torch.device = None) -> torch.Tensor:
    """[h*w, h*w] distances ||p_i - p_j||_pow between the cells of an h x w grid (row-major flattening)."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float, device=device),
                            torch.arange(w, dtype=torch.float, device=device), indexing="ij")
    pts = torch.stack((ys.flatten(), xs.flatten()), dim=1)
    return torch.cdist(pts, pts, p=pow)


_GEO_CACHE = {}


def pair_wise_point_sim(h: int, w: int, alpha: float = 1, pow: float = 2, device: torch.device = None) -> torch.Tensor:
    """Sim[i, j] = 1 / (1 + ||p_i - p_j||_pow / alpha).  The table only depends on its arguments, so it is built once
    per (h, w, alpha, pow, device) on the CPU (bit-identical to the reference's CPU result) and cached."""
    assert alpha >= 0
    key = (h, w, float(alpha), float(pow), str(device))
    table = _GEO_CACHE.get(key)
    if table is None:
        table = (1 / (1 + pair_wise_point_dist(h, w, pow, None) / alpha)).to(device)
        _GEO_CACHE[key] = table
    return table


class MyParameter(nn.Module):
    def __init__(self, shape: Iterable[int], dtype=torch.float, as_buffer: bool = False) -> None:
        super().__init__()
        self.tensor = nn.Parameter(torch.zeros(tuple(shape), dtype=dtype), requires_grad=not as_buffer)

    def _reset_parameters(self):
        nn.init.zeros_(self.tensor)

    def copy_(self, value: torch.Tensor):
        with torch.no_grad():
            self.tensor.copy_(value)

    def normalize_sum_(self, dim: int, min_val: float = 0):
        with torch.no_grad():
            normalize_sum_(self.tensor.clamp_min_(min_val), dim=dim)
