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
        """match.py:21-31.  Either the compact pair -- feat_1 [bs, D] instance embeddings, feat_2 [K, D] class embeddings --
        or the reference's own call form: the two tensors expanded to a common [bs, K, D] (match.py:72-75).  Expanded views
        (stride 0 along the broadcast axis, what `expand` / `expand_as` produce) go to the kernel on their bases, so the
        [bs, K, D] product the reference materialises is never formed; -> [bs, K]."""
        needs_grad = torch.is_grad_enabled() and (feat_1.requires_grad or feat_2.requires_grad)
        if feat_1.dim() == 2 and feat_2.dim() == 2:
            if not needs_grad:
                return native.similarity(feat_1, feat_2, self.similarity_name)
            feat_1, feat_2 = feat_1.unsqueeze(1), feat_2.unsqueeze(0)
        f1, f2 = torch.broadcast_tensors(feat_1, feat_2)
        if not needs_grad and f1.dim() == 3 and f1.stride(1) == 0 and f2.stride(0) == 0:
            return native.similarity(f1[:, 0, :], f2[0], self.similarity_name)
        # training (the kernel is forward-only) and arbitrary non-broadcast pairs: the reference's formulas on stock ops
        if self.similarity_name == "inner_product":
            return (f1 * f2).sum(-1)
        if self.similarity_name == "cosine":
            return (torch.cosine_similarity(f1, f2, dim=-1) + 1) / 2
        return 1 / (1 + torch.linalg.vector_norm(f1 - f2, dim=-1))

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
        # training (worker_schema_net.py:129-139): anything that needs a gradient goes through GNN.forward (autograd.GnnFn)
        train = torch.is_grad_enabled() and (any(p.requires_grad for p in self.gnn.parameters()) or
                                             any(t.requires_grad for t in list(vw) + list(ed)) or
                                             class_dict["class_vertices"].requires_grad or class_dict["class_edges"].requires_grad)
        if packed is not None and not train:
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
            if packed is not None:                              # lists of a packed result are views of [L]-wide slots
                n = instance_dict.sizes
                ids, vw, ed = [x[:s] for x, s in zip(ids, n)], [x[:s] for x, s in zip(vw, n)], [x[:s, :s] for x, s in zip(ed, n)]
            pid, pw, pe, mask = self._from_lists(ids, vw, ed)
            feat_instance = self.gnn(nodes=pw, edges=pe, ingredients=pid, feat_mask=mask)
            if self.mutate_inputs:
                tgt = (instance_dict["instance_ingredients"], instance_dict["instance_vertices"], instance_dict["instance_edges"])
                for i in range(len(pid)):
                    tgt[0][i], tgt[1][i], tgt[2][i] = pid[i], pw[i], pe[i]
        if not train and hasattr(class_dict, "prune_node_threshold") and class_dict["class_edges"].is_contiguous():
            # the atlas came from SchemaNet.get_atlas(): pruned vertices are known to be isolated
            feat_kg = native.gnn_forward_class(self.gnn.param_pack(), class_dict["class_vertices"], class_dict["class_edges"],
                                               class_dict["class_ingredients"], class_dict.prune_node_threshold)
        else:
            feat_kg = self.gnn(nodes=class_dict["class_vertices"], edges=class_dict["class_edges"],
                               ingredients=class_dict["class_ingredients"])
        return self.similarity(feat_instance, feat_kg)
