"""Drop-in for `schema_inference.loss.schema_inference_loss` (schema_inference/loss/schema_inference_loss.py:10-67) and the
`CELoss` it sits beside (loss/base_loss.py:16-33): cross entropy on the logits plus the entropy-sparsity regulariser on the
class atlas.  Small reductions over the atlas tensors on stock torch ops (training only; autograd provides the backward)."""
from collections import OrderedDict
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F


class Loss(nn.Module):
    def forward(self, output: Dict[str, torch.Tensor], target: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        raise NotImplementedError


def _logits(output):
    pred = output["pred"]
    return pred["pred"] if isinstance(pred, dict) else pred


class CELoss(Loss):
    def __init__(self, ignore_index: int = -100, reduction: str = "mean", **kwargs):
        super().__init__()
        self.ignore_index, self.reduction = ignore_index, reduction

    def forward(self, output, target, name: str = "cls"):
        return OrderedDict([(name, F.cross_entropy(_logits(output), target["label"], ignore_index=self.ignore_index,
                                                   reduction=self.reduction))])


def entropy(p: torch.Tensor, eps: float = 1.0e-7, dim: int = -1, keepdim: bool = False) -> torch.Tensor:
    return -(p * (p + eps).log()).sum(dim=dim, keepdim=keepdim)


def rectify_linear(x: torch.Tensor, a: float = 0):
    """x above the knee a; below it a smooth branch a - 1 + 1 / (1 + a - x) with the same value and slope at x = a."""
    return x if x > a else a - 1 + 1.0 / (1 + a - x)


class SchemaInferenceLoss(Loss):
    def __init__(self, re_a_vertex: float = 3, re_a_edge: float = 3, **kwargs):
        super().__init__()
        self.re_a_vertex, self.re_a_edge = re_a_vertex, re_a_edge

    def loss_sparsity(self, vertex_weights: torch.Tensor, edge_weights: torch.Tensor):
        ev = entropy(vertex_weights).max(dim=0)[0]                    # largest vertex entropy over the classes
        ee = entropy(edge_weights).max(dim=1)[0].mean()               # per class: largest row entropy; mean over classes
        return OrderedDict([("entropy_vertex", ev), ("entropy_edge", ee),
                            ("re_entropy_vertex", rectify_linear(ev, a=self.re_a_vertex)),
                            ("re_entropy_edge", rectify_linear(ee, a=self.re_a_edge))])

    def forward(self, output, target):
        ret = OrderedDict(cls=F.cross_entropy(_logits(output), target["label"]))
        ret.update(self.loss_sparsity(output["class_vertices"], output["class_edges"]))
        return ret
