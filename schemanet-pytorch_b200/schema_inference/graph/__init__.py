"""Drop-in for `schema_inference.graph` (schema_inference/graph/__init__.py:14-57)."""
from collections import OrderedDict

import torch
from torch import nn

from schema_inference.utils import IngredientModelWrapper
from .gnn import GNN
from .match import Matcher
from .schema_net import Atlas, InstanceGraphs, SchemaNet

__all__ = ["SchemaNet", "Matcher", "GNN", "SchemaNetPredictor", "InstanceGraphs", "Atlas"]


class SchemaNetPredictor(nn.Module):
    """backbone + discretisation (no grad) -> instance IR-graphs -> match against the class IR-atlas.

    forward(x) returns an OrderedDict: "pred" [bs, K] first, then the atlas ("class_vertices", "class_edges",
    "class_ingredients"); with requires_graph=True also the instance graphs (padded by the matcher, as in the
    reference), "ingredients" and "attn_cls".
    """

    def __init__(self, ingredient_wrapper: IngredientModelWrapper, schema_net: SchemaNet, matcher: Matcher):
        super().__init__()
        self.ingredient_wrapper = ingredient_wrapper
        self.schema_net = schema_net
        self.matcher = matcher
        self.num_classes = schema_net.num_classes

    def forward(self, x: torch.Tensor, requires_graph: bool = False):
        with torch.no_grad():
            taps = self.ingredient_wrapper(x)
        graphs = self.schema_net(ingredients=taps["ingredients"], attn=taps["attn"], attn_cls=taps["attn_cls"])
        atlas = self.schema_net.get_atlas()
        out = OrderedDict(pred=self.matcher(instance_dict=graphs, class_dict=atlas))
        out.update(atlas)
        if requires_graph:
            out.update(graphs)
            for key in ("ingredients", "attn_cls"):
                out[key] = taps[key]
        return out
