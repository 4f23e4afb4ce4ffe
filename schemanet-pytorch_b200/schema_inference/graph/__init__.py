"""Drop-in for `schema_inference.graph` (schema_inference/graph/__init__.py:14-57)."""
import collections
from typing import Dict, List

import torch
import torch.nn as nn

from .schema_net import SchemaNet, InstanceGraphs
from .match import Matcher
from .gnn import GNN

from schema_inference.utils import IngredientModelWrapper


class SchemaNetPredictor(nn.Module):
    """ingredient model -> SchemaNet instance graphs -> matcher.  Returns an OrderedDict with "pred" [bs, K], the
    atlas tensors and, with requires_graph=True, the (padded) instance graphs, "ingredients" and "attn_cls"."""

    def __init__(self, ingredient_wrapper: IngredientModelWrapper, schema_net: SchemaNet, matcher: Matcher):
        super().__init__()
        self.ingredient_wrapper = ingredient_wrapper
        self.schema_net = schema_net
        self.matcher = matcher
        self.num_classes = schema_net.num_classes

    def forward(self, x: torch.Tensor, requires_graph: bool = False):
        ret = collections.OrderedDict()
        with torch.no_grad():
            output = self.ingredient_wrapper(x)
        instance_dict: Dict[str, List[torch.Tensor]] = self.schema_net(
            ingredients=output["ingredients"], attn=output["attn"], attn_cls=output["attn_cls"])
        class_dict = self.schema_net.get_atlas()
        ret["pred"] = self.matcher(instance_dict=instance_dict, class_dict=class_dict)
        ret.update(class_dict)
        if requires_graph:
            ret.update(instance_dict)
            ret["ingredients"] = output["ingredients"]
            ret["attn_cls"] = output["attn_cls"]
        return ret
