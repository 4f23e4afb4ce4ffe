"""Atlas initialisation on the GPU: the two dataset passes of scripts/init_schema_net.py (init_class_vertices :43-65,
init_graph :19-40) with the per-sample Python accumulation loops replaced by one kernel launch per batch.

The reference iterates a DataLoader through the backbone wrapper; the backbone and the data pipeline are out of scope here
(SURVEY.md section 2), so both functions take an iterable of already-tapped batches
    {"ingredients": [bs, L] int64, "attn_cls": [bs, L], "attn": [bs, L, L], "label": [bs] int64}      (CUDA tensors)
-- exactly the fields the reference reads from `wrapper(x)` and `gt`.  The dense per-sample tensors come from
`SchemaNet.feat_to_full_vertices` / `feat_to_limited_edges` (the init-time kernels); the running sums are added in batch
order by `sh_dev_class_accumulate`, i.e. in the reference's fp32 summation order.
"""
from typing import Dict, Iterable

import torch

from . import native


@torch.no_grad()
def init_class_vertices(batches: Iterable[Dict[str, torch.Tensor]], schema_net) -> torch.Tensor:
    """scripts/init_schema_net.py:43-65 -> class_vertices [K, M], every row normalised to sum 1."""
    dev = schema_net.vertex_weights.tensor.device
    K, M = schema_net.num_classes, schema_net.num_vertices
    acc = torch.zeros(K, M, dtype=torch.float32, device=dev)
    n_tracked = torch.zeros(K, dtype=torch.float32, device=dev)
    for batch in batches:
        v = schema_net.feat_to_full_vertices(batch["ingredients"], batch["attn_cls"])          # [bs, M]
        native.class_accumulate(v, batch["label"], acc, n_tracked)
    acc /= n_tracked[:, None]
    acc /= acc.sum(dim=-1, keepdim=True)
    return acc


@torch.no_grad()
def init_graph(batches: Iterable[Dict[str, torch.Tensor]], schema_net) -> None:
    """scripts/init_schema_net.py:19-40: edge_weights[k] = mean over the samples of class k of their class-local edges, then
    `schema_net.normalize()`.  Like the reference it ADDS to whatever edge_weights holds (the caller zeroes it first)."""
    ew = schema_net.edge_weights.tensor
    K = schema_net.num_classes
    n_tracked = torch.zeros(K, dtype=torch.float32, device=ew.device)
    for batch in batches:
        e = schema_net.feat_to_limited_edges(batch["ingredients"], batch["attn"], batch["label"])   # [bs, Vc, Vc]
        native.class_accumulate(e, batch["label"], ew.data, n_tracked)
    ew.data /= n_tracked[:, None, None]
    schema_net.normalize()


@torch.no_grad()
def init_schema_net(batches, schema_net) -> None:
    """scripts/init_schema_net.py:108-124: vertex pass, keep the top-Vc codes of every class (descending weight), edge pass.
    `batches` is iterated twice (a list, or any re-iterable of tapped batches)."""
    init_weights, valid_vertices = init_class_vertices(batches, schema_net).topk(schema_net.class_max_vertices, dim=1)
    schema_net.register_class_vertices(valid_vertices)
    schema_net.vertex_weights.copy_(init_weights)
    init_graph(batches, schema_net)
