"""Training support: autograd Functions whose FORWARD is the libschemahead kernel path and whose
BACKWARD follows the reference's graph -- `SchemaNetTrainer.train_iter` (schema_inference/tasks/worker_schema_net.py:120-140)
differentiates through `SchemaNet.get_atlas` (schema_net.py:144-184, row sums detached: `normalize_sum_clamp(detach_sum=True)`)
and `GNN.forward` (gnn.py:78-98).

* atlas: closed-form backward (the normalisers are constants by `detach_sum`, so every entry only scales its own gradient);
* GNN: the backward pass recomputes the layer chain with stock torch ops under `enable_grad` and lets autograd produce the
  gradients of every input and parameter (activations are not kept between forward and backward; the value the loss sees
  is the kernels' forward result).
"""
from typing import List, Optional

import torch
import torch.nn.functional as F

from schemanet_b200 import native


class ClassVerticesFn(torch.autograd.Function):
    """class_vertices = nan_to_num(clamp_min(vw, 1e-5) / sum.detach())   (schema_net.py:144-150)."""

    @staticmethod
    def forward(ctx, vertex_weights: torch.Tensor):
        vw = vertex_weights.detach()
        cv, _ = _vertices_only(vw)
        clamped = vw.clamp_min(1.0e-5)
        ctx.save_for_backward(vw >= 1.0e-5, clamped.sum(dim=-1, keepdim=True))
        return cv

    @staticmethod
    def backward(ctx, grad_cv):
        passes, total = ctx.saved_tensors
        g = (grad_cv / total).nan_to_num(0.0, 0.0, 0.0)
        return g * passes


def _vertices_only(vw: torch.Tensor):
    K, Vc = vw.shape
    cv = torch.empty(K, Vc, dtype=torch.float32, device=vw.device)
    native.check(native.lib().sh_dev_class_atlas(native.ptr(vw.contiguous()), None, K, Vc, -1.0, 0, 0, native.ptr(cv), None,
                                                 native.stream()))
    return cv, None


class ClassEdgesFn(torch.autograd.Function):
    """class_edges (schema_net.py:152-175): prune by the vertex mask (in place on the parameter, under no_grad, and once more
    as a multiplication so that pruned entries get zero gradient), clamp_min(0), divide by the DETACHED row sums, optional
    diagonal removal.  d ce[i, j] / d ew[i, j] = keep[i, j] * [ew[i, j] >= 0] / rowsum[i]; nothing else depends on ew."""

    @staticmethod
    def forward(ctx, edge_weights: torch.Tensor, vertex_weights: torch.Tensor, prune_threshold: Optional[float],
                remove_self_loop: bool, nan_on_empty_rows: bool = True):
        ctx.nan_on_empty_rows = nan_on_empty_rows
        cv, ce = native.class_atlas(vertex_weights, edge_weights, prune_threshold, prune_in_place=True,
                                    remove_self_loop=remove_self_loop, want_edges=True)
        ew = edge_weights.detach()                              # (already pruned in place by the kernel)
        rowsum = ew.clamp_min(0).sum(dim=-1, keepdim=True)
        keep = (cv > prune_threshold) if prune_threshold is not None else torch.ones_like(cv, dtype=torch.bool)
        ctx.save_for_backward(ew, rowsum, keep)
        ctx.remove_self_loop = remove_self_loop
        return ce

    @staticmethod
    def backward(ctx, grad_ce):
        ew, rowsum, keep = ctx.saved_tensors
        # A row whose sum is 0 (a pruned vertex) is 0/0 behind nan_to_num in the reference: autograd hands every entry of such
        # a row a NaN gradient (0 * inf).  Kept by default -- parity is with the reference's live behaviour, and its
        # optimiser + normalize() turn those rows back into zeros -- unless the module opts out (zeros instead).
        g = grad_ce / rowsum
        if not ctx.nan_on_empty_rows:
            g = g.nan_to_num(0.0, 0.0, 0.0)
        g = g * (ew >= 0) * keep.unsqueeze(-1) * keep.unsqueeze(-2)
        if ctx.remove_self_loop:
            g = g.clone()
            g.diagonal(dim1=1, dim2=2).zero_()
        return g, None, None, None, None


def gnn_layers_torch(nodes, edges, ingredients, feat_mask, embedding, lin_w: List[torch.Tensor], lin_b: List[torch.Tensor],
                     ln_w: List[torch.Tensor], ln_b: List[torch.Tensor], fc_w, fc_b, eps: float):
    """gnn.py:78-98 on stock ops (used by the backward pass only)."""
    x = F.embedding(ingredients, embedding)
    D = x.shape[-1]
    adj = (edges + edges.transpose(1, 2)) / 2 + torch.eye(edges.shape[-1], device=edges.device, dtype=edges.dtype)
    for w, b, g_, b_ in zip(lin_w, lin_b, ln_w, ln_b):
        x = F.linear(torch.bmm(adj, x), w, b)
        if feat_mask is not None:
            x = x.masked_fill(feat_mask.unsqueeze(-1), 0)
        x = F.relu(F.layer_norm(x, (D,), g_, b_, eps))
    return F.linear((x * nodes.unsqueeze(-1)).mean(dim=1), fc_w, fc_b)


class GnnFn(torch.autograd.Function):
    """GNN.forward: kernels forward, recompute-and-differentiate backward."""

    @staticmethod
    def forward(ctx, gnn, feat_mask, ingredients, nodes, edges, *params):
        ctx.gnn, ctx.feat_mask, ctx.ingredients = gnn, feat_mask, ingredients
        ctx.save_for_backward(nodes, edges, *params)
        return gnn._forward_kernels(nodes.detach(), edges.detach(), ingredients, feat_mask)

    @staticmethod
    def backward(ctx, grad_out):
        nodes, edges, *params = ctx.saved_tensors
        needs = ctx.needs_input_grad[3:]
        leaves = [t.detach().requires_grad_(n) for t, n in zip([nodes, edges] + params, needs)]
        n_layers = ctx.gnn.num_layers
        emb, fc_w, fc_b = leaves[2], leaves[3], leaves[4]
        rest = leaves[5:]
        groups = [rest[i * n_layers:(i + 1) * n_layers] for i in range(4)]
        with torch.enable_grad():
            out = gnn_layers_torch(leaves[0], leaves[1], ctx.ingredients, ctx.feat_mask, emb, groups[0], groups[1], groups[2],
                                   groups[3], fc_w, fc_b, ctx.gnn.layers[0].norm.eps)
            wanted = [t for t, n in zip(leaves, needs) if n]
            grads = torch.autograd.grad(out, wanted, grad_out, allow_unused=True) if wanted else []
        it = iter(grads)
        return (None, None, None) + tuple(next(it) if n else None for n in needs)
