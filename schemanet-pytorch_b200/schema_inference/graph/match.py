"""Drop-in for `schema_inference.graph.match.Matcher` (schema_inference/graph/match.py:10-76)."""
from typing import Any, Dict, List

import torch
import torch.nn as nn

from schemanet_b200 import native
from .gnn import GNN


class Matcher(nn.Module):
    def __init__(self, similarity: str, num_codes: int, gnn_cfg: Dict[str, Any]):
        super().__init__()
        self.gnn = GNN(num_codes=num_codes, **gnn_cfg)
        if similarity not in native.SIM_KINDS:
            raise KeyError(similarity)
        self.similarity_name = similarity
        # Matcher.forward of the reference pads the caller's lists in place (match.py:49-54), which is why
        # SchemaNetPredictor(requires_graph=True) returns padded graphs; kept by default.
        self.mutate_inputs = True

    def similarity(self, feat_1: torch.Tensor, feat_2: torch.Tensor) -> torch.Tensor:
        """feat_1 [bs, D] instance embeddings vs feat_2 [K, D] class embeddings -> [bs, K]."""
        return native.similarity(feat_1, feat_2, self.similarity_name)

    def _from_lists(self, ids: List[torch.Tensor], vw: List[torch.Tensor], ed: List[torch.Tensor]):
        sizes = [len(x) for x in ids]
        bs, N = len(ids), max(sizes)
        dev = ids[0].device
        pid = torch.full((bs, N), self.gnn.num_codes, dtype=torch.int64, device=dev)
        pw = torch.zeros(bs, N, dtype=torch.float32, device=dev)
        pe = torch.zeros(bs, N, N, dtype=torch.float32, device=dev)
        for i, s in enumerate(sizes):
            pid[i, :s] = ids[i]
            pw[i, :s] = vw[i]
            pe[i, :s, :s] = ed[i]
        mask = torch.arange(N, device=dev)[None, :] >= torch.tensor(sizes, device=dev)[:, None]
        return pid, pw, pe, mask

    def forward(self, instance_dict: Dict[str, List[torch.Tensor]], class_dict: Dict[str, torch.Tensor]):
        ids = instance_dict["instance_ingredients"]      # [[n_1], ..., [n_bs]]
        vw = instance_dict["instance_vertices"]          # [[n_1], ..., [n_bs]]
        ed = instance_dict["instance_edges"]             # [[n_1, n_1], ..., [n_bs, n_bs]]
        packed = getattr(instance_dict, "packed", None)
        if packed is not None:
            feat_instance = self.gnn.forward_packed(packed)
            if self.mutate_inputs:
                N, L, bs = max(instance_dict.sizes), packed.L, packed.B
                live = torch.arange(N, device=packed.ids.device)[None, :] < packed.num_vertices[:, None]
                pid = torch.where(live, packed.ids[:, :N], self.gnn.num_codes)
                pw = torch.where(live, packed.vertex_w[:, :N], 0.0)
                pe = packed.edges.view(bs, L, L)[:, :N, :N]     # zero-padded in place by stage 2 (SH_G_ZERO_PAD)
                for i in range(bs):
                    ids[i], vw[i], ed[i] = pid[i], pw[i], pe[i]
        else:
            pid, pw, pe, mask = self._from_lists(ids, vw, ed)
            feat_instance = self.gnn(nodes=pw, edges=pe, ingredients=pid, feat_mask=mask)
            if self.mutate_inputs:
                for i in range(len(ids)):
                    ids[i], vw[i], ed[i] = pid[i], pw[i], pe[i]
        if hasattr(class_dict, "prune_node_threshold") and class_dict["class_edges"].is_contiguous():
            # the atlas came from SchemaNet.get_atlas(): pruned vertices are known to be isolated
            self.gnn._check_inference()
            feat_kg = native.gnn_forward_class(self.gnn.param_pack(), class_dict["class_vertices"], class_dict["class_edges"],
                                               class_dict["class_ingredients"], class_dict.prune_node_threshold)
        else:
            feat_kg = self.gnn(nodes=class_dict["class_vertices"], edges=class_dict["class_edges"],
                               ingredients=class_dict["class_ingredients"])
        return self.similarity(feat_instance, feat_kg)
