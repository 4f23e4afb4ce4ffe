"""Drop-in for `schema_inference.utils.ingredient_model_wrapper.IngredientModelWrapper`
(schema_inference/utils/ingredient_model_wrapper.py:9-69).

The backbone stays the caller's (JIT) module and is out of kernel scope.  Everything after it runs in libschemahead:
the nearest-codeword search uses the codebook taken from `discretization_jit.discretization.vocabulary.weight`
(the reference registers the same tensor as `discretization_tensor`, :28) instead of calling the traced
cdist + argmin graph, and the head-mean + slicing of the raw attention is one kernel.
"""
import collections
from typing import Dict

import torch
import torch.nn as nn

from schemanet_b200 import native


class IngredientModelWrapper(nn.Module):
    """Always works in evaluation mode.  Returns
        cls_token [bs, 1, dim], feat [bs, L, dim], feat_origin [bs, L, dim], ingredients [bs, L] (int64),
        attn [bs, L, L], attn_cls [bs, L]   (attention = raw logits averaged over heads)
    Set `full_outputs = False` to skip the three tensors the predictor never reads (cls_token, feat, feat_origin)."""

    def __init__(self, backbone_jit, discretization_jit=None):
        super().__init__()
        self.backbone_jit = backbone_jit
        self.discretization_jit = discretization_jit
        self.register_buffer("discretization_tensor", discretization_jit.discretization.vocabulary.weight)
        self.num_ingredients: int = self.discretization_tensor.shape[0]
        self.emb_dim: int = self.discretization_tensor.shape[1]
        self.full_outputs = True
        self.kernel_mode = native.DISC_AUTO

    def train(self, mode: bool = True):
        self.training = mode
        for module in self.children():
            module.train(False)
        return self

    def eval(self):
        return self.train(False)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        ret: Dict[str, torch.Tensor] = collections.OrderedDict()
        out_backbone = self.backbone_jit(x)
        mid_feat: torch.Tensor = out_backbone["mid_feat"].contiguous()          # [1 + L, bs, dim], sequence first
        extracted_attn = out_backbone["extracted"] if "extracted" in out_backbone else None
        T, bs, dim = mid_feat.shape
        L = T - 1
        vocab = self.discretization_tensor
        ingredients = torch.empty(bs, L, dtype=torch.int64, device=mid_feat.device)
        gathered = torch.empty(L * bs, dim, dtype=torch.float32, device=mid_feat.device) if self.full_outputs else None
        # token-major rows (r = t*bs + b) written straight into the [bs, L] layout the head consumes
        native.discretize(mid_feat[1:].reshape(L * bs, dim), vocab, out_idx=ingredients, idx_rows=bs, idx_row_stride=L,
                          idx_col_stride=1, out_seq=gathered, mode=self.kernel_mode)
        if self.full_outputs:
            ret["cls_token"] = mid_feat[:1].transpose(0, 1).contiguous()
            ret["feat"] = gathered.view(L, bs, dim).transpose(0, 1).contiguous()
            ret["feat_origin"] = mid_feat[1:].transpose(0, 1).contiguous()
        ret["ingredients"] = ingredients
        if extracted_attn is not None:
            ret["attn"], ret["attn_cls"] = native.attention_prologue(extracted_attn, bs)
        else:
            ret["attn"] = torch.zeros(bs, L, L, device=x.device)
            ret["attn_cls"] = torch.zeros(bs, L, device=x.device)
        return ret
