"""Drop-in for `discretization.visual_word_encoder` (discretization/visual_word_encoder.py:10-68) and for the
`DiscretizationJitWrapper` of scripts/save_backbone_jit.py:121-131 (the module the reference traces into
`discretization-jit.pth`)."""
from typing import Dict

import torch
import torch.nn as nn
from torch.utils.hooks import RemovableHandle

from .discretization import Discretization


class Adapter:
    """Strips the cls token before discretization and puts it back afterwards (ViT: one cls token)."""

    def __init__(self):
        self.shape: torch.Size = None

    def adapt(self, x: torch.Tensor) -> torch.Tensor:
        self.cls_token = x[:1]
        return x[1:]

    def reconstruct(self, x: torch.Tensor, match: torch.Tensor) -> torch.Tensor:
        return torch.cat((self.cls_token, x), dim=0), match


class DiscretizationJitWrapper(nn.Module):
    """mid_feat [1+L, bs, d] -> (sequence with the patch tokens replaced by their codewords, ingredients [L, bs])."""

    def __init__(self, discretization: Discretization):
        super().__init__()
        self.discretization = discretization
        self.adapter = Adapter()

    def forward(self, dummy_input: torch.Tensor):
        seq = self.adapter.adapt(dummy_input)
        output, match = self.discretization(seq)
        return self.adapter.reconstruct(output, match)


class VisualWordEncoder:
    def __init__(self, model: nn.Module, encode_layer: str, discretization: Discretization):
        self.encode_layer = encode_layer
        self.discretization = discretization
        self.adapter = Adapter()
        self.hook = self.register_forward_hooks(model)
        self.mid_dict: Dict[str, torch.Tensor] = {"origin_seq": None, "encoded_seq": None, "match": None}

    def register_forward_hooks(self, model: nn.Module) -> RemovableHandle:
        raw_model = model.module if isinstance(model, nn.parallel.DistributedDataParallel) else model
        for name, module in raw_model.named_modules():
            if name == self.encode_layer:
                def forward_hook(module, input, output):
                    self.mid_dict["origin_seq"] = output
                    seq, match = self.discretization(self.adapter.adapt(output))
                    seq, match = self.adapter.reconstruct(seq, match)
                    self.mid_dict["encoded_seq"] = seq
                    self.mid_dict["match"] = match
                    return seq
                return module.register_forward_hook(forward_hook)

    def clear(self):
        self.hook.remove()
        for k in self.mid_dict:
            self.mid_dict[k] = None
