"""Drop-in for `discretization.visual_word_encoder` (discretization/visual_word_encoder.py:10-68) and for the
`DiscretizationJitWrapper` of scripts/save_backbone_jit.py:121-131 (the module the reference traces into
`discretization-jit.pth`)."""
from typing import Dict, Optional

import torch
from torch import nn
from torch.utils.hooks import RemovableHandle

from .discretization import Discretization


class Adapter:
    """ViT sequences carry one cls token in front: `adapt` strips it, `reconstruct` puts it back."""

    def __init__(self):
        self.shape: torch.Size = None
        self.cls_token: Optional[torch.Tensor] = None

    def adapt(self, x: torch.Tensor) -> torch.Tensor:
        self.cls_token, patches = x[:1], x[1:]
        return patches

    def reconstruct(self, x: torch.Tensor, match: torch.Tensor):
        return torch.cat((self.cls_token, x), dim=0), match


class DiscretizationJitWrapper(nn.Module):
    """mid_feat [1+L, bs, d] -> (sequence with the patch tokens replaced by their codewords, ingredients [L, bs])."""

    def __init__(self, discretization: Discretization):
        super().__init__()
        self.discretization = discretization
        self.adapter = Adapter()

    def forward(self, dummy_input: torch.Tensor):
        patches = self.adapter.adapt(dummy_input)
        encoded, match = self.discretization(patches)
        return self.adapter.reconstruct(encoded, match)


class VisualWordEncoder:
    """Forward hook that discretises the output of `encode_layer` inside a backbone and records what it did in
    `mid_dict` ("origin_seq", "encoded_seq", "match")."""

    KEYS = ("origin_seq", "encoded_seq", "match")

    def __init__(self, model: nn.Module, encode_layer: str, discretization: Discretization):
        self.encode_layer = encode_layer
        self.discretization = discretization
        self.adapter = Adapter()
        self.mid_dict: Dict[str, torch.Tensor] = dict.fromkeys(self.KEYS)
        self.hook = self.register_forward_hooks(model)

    def _on_forward(self, module, inputs, output):
        encoded, match = self.discretization(self.adapter.adapt(output))
        encoded, match = self.adapter.reconstruct(encoded, match)
        self.mid_dict.update(origin_seq=output, encoded_seq=encoded, match=match)
        return encoded

    def register_forward_hooks(self, model: nn.Module) -> RemovableHandle:
        if isinstance(model, nn.parallel.DistributedDataParallel):
            model = model.module
        target = dict(model.named_modules()).get(self.encode_layer)
        return None if target is None else target.register_forward_hook(self._on_forward)

    def clear(self):
        self.hook.remove()
        self.mid_dict = dict.fromkeys(self.KEYS)
