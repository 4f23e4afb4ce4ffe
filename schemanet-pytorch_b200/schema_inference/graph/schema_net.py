"""Drop-in for `schema_inference.graph.schema_net.SchemaNet` (schema_inference/graph/schema_net.py:10-399).

Same constructor keywords, parameters / state-dict keys, methods and return structures.  What changed is where the
work happens: the reference soft-maxes on the device, ships everything to the CPU, loops in single-threaded C++ and
ships B small tensors back; here `forward` is ONE kernel launch that reads the raw attention
once (clamp + softmax + code ranking + block means + normalisation + attribute mix fused, `sh_dev_instance_graphs`)
and `get_atlas` is two launches over the class tensors (`sh_dev_class_atlas`).  The only host synchronisation left is
the single D2H copy of the B graph sizes that a list-returning API needs.
"""
import logging
from typing import Tuple, List, Dict

import torch
import torch.nn as nn

import schema_inference.graph.utils as graph_utils
from schemanet_b200 import native


class InstanceGraphs(dict):
    """The dict `SchemaNet.forward` returns ({"instance_ingredients", "instance_vertices", "instance_edges"} -> lists
    of per-image tensors), plus a handle on the packed device buffers the lists are views of, so that `Matcher` can
    consume them without re-packing."""
    packed: "native.PackedGraphs" = None
    sizes: List[int] = None


class Atlas(dict):
    """The dict `SchemaNet.get_atlas` returns, plus the prune threshold the class graphs were built with: vertices at
    or below it have all-zero edge rows/columns, which lets `Matcher` skip them (same result, fewer flops)."""
    prune_node_threshold: float = None


class SchemaNet(nn.Module):
    """IR-Atlas (class graphs) and instance IR-Graph generation.

    Parameters: `vertex_weights.tensor` [K, Vc], `edge_weights.tensor` [K, Vc, Vc], `vertex_attribute_weights.tensor`
    and `edge_attribute_weights.tensor` [2, 1]; buffer-like `class_ingredients.tensor` [K, Vc] (int64).
    """

    def __init__(
        self,
        num_vertices: int,
        num_classes: int = 10,
        dist_alpha: float = 1,
        dist_pow: float = 2,
        feat_h: int = 14,
        feat_w: int = 14,
        class_max_vertices: int = None,
        constant_vertex_attr: Tuple[float, float] = None,
        constant_edge_attr: Tuple[float, float] = None,
        clamp_vertex_attn: float = None,
        clamp_edge_attn: float = None,
        remove_self_loop: bool = False,
        prune_node_threshold: float = None,
        apply_normalize: bool = True,
        clamp_weights: bool = True
    ):
        super().__init__()
        self.logger = logging.getLogger("SchemaNet")
        cfg = dict(locals())
        for name in ("num_vertices", "num_classes", "dist_alpha", "dist_pow", "feat_h", "feat_w", "constant_vertex_attr",
                     "constant_edge_attr", "clamp_vertex_attn", "clamp_edge_attn", "remove_self_loop",
                     "prune_node_threshold", "apply_normalize", "clamp_weights"):
            setattr(self, name, cfg[name])
        # the reference also writes -inf into the caller's attention tensors (schema_net.py:296,335); kept by default
        self.write_back_clamp = True
        # training: the reference's autograd gives fully pruned edge rows NaN gradients (0/0 behind nan_to_num); kept by default
        self.nan_grad_on_pruned_rows = True

        self.register_buffer("n_tracked", torch.zeros(num_classes), persistent=False)
        if class_max_vertices is None:
            class_max_vertices = num_vertices
        assert class_max_vertices <= num_vertices
        self.class_max_vertices = class_max_vertices

        P = graph_utils.MyParameter
        self.class_ingredients = P((num_classes, class_max_vertices), dtype=torch.long, as_buffer=True)
        self.class_ingredient_dict: List[Dict[int, int]] = list()
        self.vertex_weights = P((num_classes, class_max_vertices))
        self.edge_weights = P((num_classes, class_max_vertices, class_max_vertices))
        self.vertex_attribute_weights = P((2, 1), as_buffer=constant_vertex_attr is not None)
        self.edge_attribute_weights = P((2, 1), as_buffer=constant_edge_attr is not None)
        self._reset_parameters()

    # ------------------------------------------------------------------ parameters
    def _reset_parameters(self):
        for p, const in ((self.vertex_attribute_weights, self.constant_vertex_attr),
                         (self.edge_attribute_weights, self.constant_edge_attr)):
            nn.init.constant_(p.tensor, 0.5)
            if const is not None:
                p.copy_(torch.tensor(const).reshape(2, 1))
        for p in (self.vertex_weights, self.edge_weights):
            nn.init.trunc_normal_(p.tensor, mean=0.5, std=1 / 6, a=0, b=1)
            p.normalize_sum_(dim=-1)
        self.normalize()

    def register_class_vertices(self, class_vertices: torch.LongTensor):
        self.class_ingredients.copy_(class_vertices)
        self.class_ingredient_dict.clear()
        for row in class_vertices.tolist():
            self.class_ingredient_dict.append({code: slot for slot, code in enumerate(row)})

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        ret = super().load_state_dict(state_dict, strict)
        self.register_class_vertices(self.class_ingredients.tensor)
        return ret

    @torch.no_grad()
    def normalize(self):
        if self.clamp_weights:
            self.vertex_attribute_weights.tensor.clamp_(min=0.01, max=10)
            self.edge_attribute_weights.tensor.clamp_(min=0.01, max=10)
        if self.apply_normalize:
            self.vertex_weights.normalize_sum_(dim=-1)
            self.edge_weights.normalize_sum_(dim=-1)
            if self.remove_self_loop:
                self.edge_weights.tensor.diagonal(dim1=1, dim2=2).fill_(0)

    # ------------------------------------------------------------------ class atlas (schema_net.py:144-184)
    def _atlas(self, want_edges: bool):
        return native.class_atlas(self.vertex_weights.tensor, self.edge_weights.tensor, self.prune_node_threshold,
                                  prune_in_place=True, remove_self_loop=self.remove_self_loop, want_edges=want_edges)

    def _grad_mode(self) -> bool:
        return torch.is_grad_enabled() and (self.vertex_weights.tensor.requires_grad or self.edge_weights.tensor.requires_grad)

    def get_class_vertices(self, detach: bool = False) -> torch.Tensor:
        if not detach and self._grad_mode():
            from .autograd import ClassVerticesFn
            return ClassVerticesFn.apply(self.vertex_weights.tensor)
        return self._atlas(False)[0]

    def get_class_edges(self, detach: bool = False) -> torch.Tensor:
        if not detach and self._grad_mode():
            from .autograd import ClassEdgesFn
            return ClassEdgesFn.apply(self.edge_weights.tensor, self.vertex_weights.tensor.detach(), self.prune_node_threshold,
                                      self.remove_self_loop, self.nan_grad_on_pruned_rows)
        return self._atlas(True)[1]

    def get_atlas(self, detach: bool = False) -> Dict[str, torch.Tensor]:
        if not detach and self._grad_mode():
            # training (worker_schema_net.py:129-139): the atlas tensors carry autograd history to the two parameters
            atlas = Atlas(class_vertices=self.get_class_vertices(), class_edges=self.get_class_edges(),
                          class_ingredients=self.class_ingredients.tensor)
            atlas.prune_node_threshold = self.prune_node_threshold
            return atlas
        class_vertices, class_edges = self._atlas(True)
        atlas = Atlas(class_vertices=class_vertices, class_edges=class_edges,
                      class_ingredients=self.class_ingredients.tensor)
        atlas.prune_node_threshold = self.prune_node_threshold
        return atlas

    def _geo(self, device) -> torch.Tensor:
        return graph_utils.pair_wise_point_sim(self.feat_h, self.feat_w, self.dist_alpha, self.dist_pow, device)

    # ------------------------------------------------------------------ initialisation-time dense variants
    def feat_to_full_vertices(self, ingredients: torch.LongTensor, attn_cls: torch.Tensor) -> torch.Tensor:
        """[bs, L] codes + raw cls attention -> vertex weights of every sample over the whole vocabulary [bs, M]."""
        if self.clamp_vertex_attn is not None:
            attn_cls.masked_fill_(attn_cls < self.clamp_vertex_attn, float("-inf"))
        attrs = self._feat_to_full_v(ingredients, attn_cls.softmax(dim=-1))
        graph_utils.normalize_max_(attrs, dim=1)
        return (attrs @ self.vertex_attribute_weights.tensor).squeeze_(-1)

    def _feat_to_full_v(self, ingredients: torch.LongTensor, attn_cls: torch.Tensor) -> torch.Tensor:
        from cpp_extension import cpp_feat_to_v_attr
        return cpp_feat_to_v_attr(ingredients, attn_cls, n_vertices=self.num_vertices, mean=True)

    def feat_to_limited_edges(self, ingredients: torch.LongTensor, attn: torch.Tensor, label: torch.LongTensor) -> torch.Tensor:
        """-> [bs, Vc, Vc] edges arranged by the class-local order of `class_ingredients[label]`."""
        if self.clamp_edge_attn is not None:
            attn.masked_fill_(attn < self.clamp_edge_attn, float("-inf"))
        attrs = self._feat_to_e(ingredients, torch.softmax(attn, dim=-1), self._geo(ingredients.device), label)
        graph_utils.normalize_sum_(attrs, dim=2)
        if self.remove_self_loop:
            attrs.diagonal(dim1=1, dim2=2).fill_(0)
        return (attrs @ self.edge_attribute_weights.tensor).squeeze_(-1)

    def _feat_to_e(self, ingredients, attn, geo_sim, label) -> torch.Tensor:
        assert len(self.class_ingredient_dict) > 0, "run `register_class_vertices` before"
        dev = ingredients.device
        ci = self.class_ingredients.tensor.to(dev)
        if label.numel() and (int(label.min()) < 0 or int(label.max()) >= self.num_classes):
            raise IndexError(f"label out of range [0, {self.num_classes})")      # (the kernels index class rows unchecked)
        if ingredients.is_cuda:
            return native.feat_to_e(ingredients, attn, geo_sim, ci, label.to(dev), self.class_max_vertices, True)
        return native.host_feat_to_e(ingredients, attn, geo_sim, ci, label, self.class_max_vertices, True)

    # ------------------------------------------------------------------ prediction path
    def _attr_grad(self) -> bool:
        return torch.is_grad_enabled() and (self.vertex_attribute_weights.tensor.requires_grad or
                                            self.edge_attribute_weights.tensor.requires_grad)

    def _build(self, ingredients, attn, attn_cls, want_vertices=True, want_edges=True, w_v=None, w_e=None) -> "native.PackedGraphs":
        if want_edges and self.remove_self_loop:
            # the reference cannot build instance edges with this flag: `instance_edges.diagonal(0, 1)` asks for
            # dim1 == dim2 and throws (cpp_extension/src/large_scale_feat_to_e.cpp:136-139); same error here
            raise RuntimeError("diagonal dimensions cannot be identical 1, 1")
        dev = ingredients.device
        return native.instance_graphs(
            ingredients, attn, attn_cls, self._geo(dev) if want_edges else None,
            (self.vertex_attribute_weights.tensor if w_v is None else w_v) if want_vertices else None,
            (self.edge_attribute_weights.tensor if w_e is None else w_e) if want_edges else None,
            clamp_vertex=self.clamp_vertex_attn, clamp_edge=self.clamp_edge_attn, raw_logits=True, mean=True,
            write_back_clamp=self.write_back_clamp, zero_pad=want_edges, want_vertices=want_vertices,
            want_edges=want_edges)

    def feat_to_instance_vertices(self, ingredients: torch.LongTensor, attn_cls: torch.Tensor
                                  ) -> Tuple[List[torch.LongTensor], List[torch.Tensor]]:
        """[bs, L] codes + raw cls attention -> (ids of every image [n_i], vertex weights of every image [n_i])."""
        g = self._build(ingredients, None, attn_cls, want_edges=False)
        ids, vw, _, _ = g.to_lists()
        return ids, vw

    def feat_to_instance_edges(self, ingredients: torch.LongTensor, attn: torch.Tensor,
                               instance_ingredients: List[torch.LongTensor]) -> List[torch.Tensor]:
        """[bs, L] codes + raw attention [bs, L, L] -> list of edge matrices [n_i, n_i] in the order of
        `instance_ingredients` (always the sorted distinct codes of each image)."""
        g = self._build(ingredients, attn, None, want_vertices=False)
        return g.to_lists()[2]

    def forward(self, ingredients: torch.LongTensor, attn: torch.Tensor, attn_cls: torch.Tensor
                ) -> Dict[str, List[torch.Tensor]]:
        """ingredients [bs, L] int64, attn [bs, L, L] and attn_cls [bs, L] RAW attention logits ->
        {"instance_ingredients": B x [n_i], "instance_vertices": B x [n_i], "instance_edges": B x [n_i, n_i]}."""
        out = InstanceGraphs()
        if self._attr_grad():
            # training of the two [2,1] attribute weights: produce both normalised channels with unit weights and
            # leave the 2 -> 1 mix to autograd, as the reference's trailing matmul does
            dev = ingredients.device
            e0, e1 = torch.tensor([1.0, 0.0], device=dev), torch.tensor([0.0, 1.0], device=dev)
            wb, self.write_back_clamp = self.write_back_clamp, False
            g0 = self._build(ingredients, attn, attn_cls, w_v=e0, w_e=e0)
            self.write_back_clamp = wb
            g1 = self._build(ingredients, attn, attn_cls, w_v=e1, w_e=e1)
            Wv, We = self.vertex_attribute_weights.tensor, self.edge_attribute_weights.tensor
            vw = g0.vertex_w * Wv[0, 0] + g1.vertex_w * Wv[1, 0]
            ed = g0.edges * We[0, 0] + g1.edges * We[1, 0]
            n = g0.num_vertices.tolist()
            B, L = ingredients.shape
            out["instance_ingredients"] = [g0.ids[b, :n[b]] for b in range(B)]
            out["instance_vertices"] = [vw[b, :n[b]] for b in range(B)]
            out["instance_edges"] = [ed[b].view(L, L)[:n[b], :n[b]] for b in range(B)]
            return out
        g = self._build(ingredients, attn, attn_cls)
        ids, vw, ed, n = g.to_lists()
        out["instance_ingredients"], out["instance_vertices"], out["instance_edges"] = ids, vw, ed
        out.packed, out.sizes = g, n
        return out
