"""Drop-in for the reference's `discretization` package: the same four public names, backed by libschemahead."""
from typing import Any, Dict

from torch import nn

from .discretization import Discretization
from .visual_word_encoder import Adapter, DiscretizationJitWrapper, VisualWordEncoder

__all__ = ["Discretization", "VisualWordEncoder", "Adapter", "DiscretizationJitWrapper", "get_visual_word_encoder"]


def get_visual_word_encoder(discretization_cfg: Dict[str, Any], model: nn.Module) -> VisualWordEncoder:
    """Builds the codebook from cfg["vocabulary"] and hooks it behind the layer named cfg["encoder_layer"]."""
    codebook = Discretization(**discretization_cfg["vocabulary"])
    layer_name = discretization_cfg["encoder_layer"]
    return VisualWordEncoder(model, layer_name, codebook)
