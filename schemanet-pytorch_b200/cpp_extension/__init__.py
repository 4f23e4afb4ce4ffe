"""Drop-in for the reference's `cpp_extension` package (cpp_extension/__init__.py:5-76, src/extension.cpp:6-12).

Same four functions, same positional signatures, same return conventions -- executed by libschemahead's CUDA kernels
instead of the reference's single-threaded CPU loops:

  * CPU tensors (what the reference's callers pass, schema_net.py:314-315,367-369) go through the library's
    host-buffer entry points (`sh_host_*`): H2D copy, kernel, D2H copy, synchronous -- the reference contract.
  * CUDA tensors skip the copies and stay on the device (`sh_dev_*`).

Outputs follow the reference: instance results live on the device of the attribute-weight tensor, `num_vertices`
is a CPU int64 tensor, dense results live where the inputs live.  When the attribute weights require grad, the final
2->1 mix is left to autograd exactly like the reference's trailing `matmul` (large_scale_feat_to_v.cpp:125,
large_scale_feat_to_e.cpp:140).
"""
from typing import Dict, List

import torch

from schemanet_b200 import native

__all__ = [
    "cpp_feat_to_v_attr",
    "cpp_feat_to_instance_v",
    "cpp_feat_to_e"
]


def _needs_grad(w: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and w.requires_grad


def _unit(device, which):
    return torch.tensor([1.0, 0.0] if which == 0 else [0.0, 1.0], device=device)


def _lengths_mask(nv: torch.Tensor, L: int) -> torch.Tensor:
    return torch.arange(L, device=nv.device)[None, :] < nv[:, None]


def feat_to_v_attr(ingredients, attn_cls, n_vertices, mean=False, ingredients_only=False):
    """ext::feat_to_v_attr (feat_to_v_attr.cpp:74-148) -> [B, n_vertices, 2]."""
    if ingredients.is_cuda:
        return native.feat_to_v_attr(ingredients, attn_cls, int(n_vertices), mean, ingredients_only)
    return native.host_feat_to_v_attr(ingredients, attn_cls, int(n_vertices), mean, ingredients_only)


def _instance_v_raw(ingredients, attn_cls, w2, mean):
    """-> (ids [B,L] slots, w [B,L] slots, nv int64 [B]) on the inputs' device."""
    if ingredients.is_cuda:
        g = native.instance_graphs(ingredients, None, attn_cls.contiguous(), None, w2.to(ingredients.device), None,
                                   raw_logits=False, mean=mean, want_edges=False)
        return g.ids, g.vertex_w, g.num_vertices.long()
    return native.host_feat_to_instance_v(ingredients, attn_cls, w2, mean)


def feat_to_instance_v(ingredients, attn_cls, vertex_attribute_weights, mean=False):
    """ext::feat_to_instance_v (large_scale_feat_to_v.cpp:41-143)
    -> [cat ids (int64), cat vertex weights (fp32), num_vertices (int64, CPU)]."""
    W = vertex_attribute_weights
    dev = W.device
    L = ingredients.shape[1]
    if _needs_grad(W):
        ids, a0, nv = _instance_v_raw(ingredients, attn_cls, _unit(ingredients.device, 0), mean)
        _, a1, _ = _instance_v_raw(ingredients, attn_cls, _unit(ingredients.device, 1), mean)
        mask = _lengths_mask(nv, L)
        attrs = torch.stack((a0[mask], a1[mask]), dim=-1).to(dev)
        w = attrs.matmul(W).squeeze(-1)
    else:
        ids, vw, nv = _instance_v_raw(ingredients, attn_cls, W.detach().reshape(-1), mean)
        mask = _lengths_mask(nv, L)
        w = vw[mask].to(dev)
    return [ids[mask].to(dev), w, nv.cpu()]


def feat_to_e(ingredients, attn, geo_sim, class_ingredient_dict, label, n_max, mean=False):
    """ext::feat_to_e (feat_to_e.cpp:31-127) -> [B, n_max, n_max, 2].
    `class_ingredient_dict` is the reference's list (len K) of {code: class-local index}."""
    K = len(class_ingredient_dict)
    table = torch.full((K, int(n_max)), -1, dtype=torch.int64)
    for k, d in enumerate(class_ingredient_dict):
        if len(d):
            codes = torch.tensor(list(d.keys()), dtype=torch.int64)
            slots = torch.tensor(list(d.values()), dtype=torch.int64)
            table[k, slots] = codes
    label_t = torch.as_tensor(label, dtype=torch.int64).reshape(-1)
    if label_t.numel() and (int(label_t.min()) < 0 or int(label_t.max()) >= K):
        raise IndexError("label out of range of class_ingredient_dict")
    if ingredients.is_cuda:
        dev = ingredients.device
        return native.feat_to_e(ingredients, attn, geo_sim.to(dev), table.to(dev), label_t.to(dev), int(n_max), mean)
    return native.host_feat_to_e(ingredients, attn, geo_sim, table, label_t, int(n_max), mean)


def _canonical(d: Dict[int, int]) -> bool:
    """True if the dictionary is {sorted distinct code: rank} -- what schema_net.py:345-348 always builds."""
    prev = None
    for i, (k, v) in enumerate(d.items()):
        if v != i or (prev is not None and k <= prev):
            return False
        prev = k
    return True


def _instance_e_raw(ingredients, attn, geo_sim, w2, mean):
    """-> (edges [B, L, L] slots, nv int64 [B]) on the inputs' device."""
    B, L = ingredients.shape
    if ingredients.is_cuda:
        dev = ingredients.device
        g = native.instance_graphs(ingredients, attn.contiguous(), None, geo_sim.to(dev), None, w2.to(dev),
                                   raw_logits=False, mean=mean, want_vertices=False)
        return g.edges.view(B, L, L), g.num_vertices.long()
    e, nv = native.host_feat_to_instance_e(ingredients, attn, geo_sim, w2, mean)
    return e.view(B, L, L), nv


def feat_to_instance_e(ingredients, attn, geo_sim, batch_ingredient_dict, edge_attribute_weights, mean=False,
                       remove_self_loop=False):
    """ext::feat_to_instance_e (large_scale_feat_to_e.cpp:33-150) -> list of B tensors [n_i, n_i] on W's device."""
    B, L = ingredients.shape
    if len(batch_ingredient_dict) != B:
        raise RuntimeError("Batch size is not compat with `batch_ingredient_dict`")   # large_scale_feat_to_e.cpp:53-56
    if remove_self_loop:
        # the reference calls diagonal(0, 1) == diagonal(offset=0, dim1=1, dim2=1) and throws
        raise RuntimeError("diagonal dimensions cannot be identical 1, 1")
    W = edge_attribute_weights
    dev = W.device
    if _needs_grad(W):
        e0, nv = _instance_e_raw(ingredients, attn, geo_sim, _unit(ingredients.device, 0), mean)
        e1, _ = _instance_e_raw(ingredients, attn, geo_sim, _unit(ingredients.device, 1), mean)
        Wd = W.to(e0.device)
        slots = e0 * Wd[0, 0] + e1 * Wd[1, 0]
    else:
        slots, nv = _instance_e_raw(ingredients, attn, geo_sim, W.detach().reshape(-1), mean)
    n = nv.tolist()
    out = []
    for b, d in enumerate(batch_ingredient_dict):
        e = slots[b, :n[b], :n[b]]
        if len(d) != n[b] or not _canonical(d):
            # arbitrary code -> index dictionary: scatter the sorted-rank result to the caller's indices
            codes = sorted(set(ingredients[b].tolist()))
            index = torch.tensor([d[c] for c in codes], dtype=torch.int64, device=e.device)
            full = torch.zeros(len(d), len(d), dtype=e.dtype, device=e.device)
            full[index[:, None], index[None, :]] = e
            e = full
        out.append(e.to(dev))
    return out


def cpp_feat_to_v_attr(
    ingredients: torch.LongTensor,
    attn_cls: torch.Tensor,
    n_vertices: int,
    mean: bool = False,
    ingredients_only: bool = False
) -> torch.Tensor:
    return feat_to_v_attr(ingredients, attn_cls, n_vertices, mean, ingredients_only)


def cpp_feat_to_instance_v(
    ingredients: torch.LongTensor,
    attn_cls: torch.Tensor,
    vertex_attribute_weights: torch.Tensor,
    mean: bool = False
) -> List[torch.Tensor]:
    return feat_to_instance_v(ingredients, attn_cls, vertex_attribute_weights, mean)


def cpp_feat_to_e(
    ingredients: torch.LongTensor,
    attn: torch.Tensor,
    geo_sim: torch.Tensor,
    class_ingredient_dict: List[Dict[int, int]],
    label: List[int],
    n_max: int,
    mean: bool = False
) -> torch.Tensor:
    return feat_to_e(ingredients, attn, geo_sim, class_ingredient_dict, label, n_max, mean)


def cpp_feat_to_instance_e(
    ingredients: torch.LongTensor,
    attn: torch.Tensor,
    geo_sim: torch.Tensor,
    batch_ingredient_dict: List[Dict[int, int]],
    edge_attribute_weights: torch.Tensor,
    mean: bool = False,
    remove_self_loop: bool = False
) -> List[torch.Tensor]:
    return feat_to_instance_e(ingredients, attn, geo_sim, batch_ingredient_dict, edge_attribute_weights, mean,
                              remove_self_loop)
