"""Drop-in for `schema_inference.graph.gnn` (schema_inference/graph/gnn.py:7-98): GraphConv / Layer / GNN with the
reference's parameter names (`embedding.weight`, `layers.{i}.g_conv.linear.*`, `layers.{i}.norm.*`, `fc.*`).

`GNN.forward` runs in libschemahead (`sh_dev_gnn_forward`): the symmetrised adjacency (E + E^T)/2 + I is formed while
tiles are staged (the reference materialises three [bs, n, n] temporaries per layer), the embedding gather is fused
into the first product, bias into the GEMM epilogue, LayerNorm + ReLU (+ the weighted mean pooling on the last layer)
into one pass.  Training: `GNN.forward` in grad mode goes through `autograd.GnnFn` (kernels forward, recompute-and-differentiate
backward); the fused packed / class-side entry points stay forward-only.
"""
import torch
import torch.nn as nn

from schemanet_b200 import native

_ACTIVATIONS = {"relu": nn.ReLU, "gelu": nn.GELU, "glu": nn.GLU, "swish": nn.SiLU, "sigmoid": nn.Sigmoid,
                "hard_sigmoid": nn.Hardsigmoid, "none": nn.Identity}


class GraphConv(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, identity_proj: bool = False):
        super().__init__()
        if identity_proj:
            assert in_dim == out_dim
        self.linear = nn.Identity() if identity_proj else nn.Linear(in_dim, out_dim)
        if isinstance(self.linear, nn.Linear):
            nn.init.xavier_uniform_(self.linear.weight)
            nn.init.normal_(self.linear.bias)


class Layer(nn.Module):
    def __init__(self, emb_dim: int, activation: str, identity_proj: bool = False):
        super().__init__()
        self.g_conv = GraphConv(emb_dim, emb_dim, identity_proj)
        self.norm = nn.LayerNorm(emb_dim)
        self.activation = _ACTIVATIONS[activation]()


class GNN(nn.Module):
    def __init__(self, num_codes: int, embed_dim: int, num_layers: int, identity_proj: bool = False,
                 activation: str = "relu"):
        super().__init__()
        self.num_codes = num_codes
        self.embed_dim = embed_dim
        self.num_layers = num_layers
        self.identity_proj = identity_proj
        self.activation_name = activation
        self.embedding = nn.Embedding(num_embeddings=num_codes + 1, embedding_dim=embed_dim, padding_idx=num_codes)
        self.layers = nn.ModuleList([Layer(embed_dim, activation, identity_proj) for _ in range(num_layers)])
        self.fc = nn.Linear(embed_dim, embed_dim)
        nn.init.normal_(self.fc.weight)
        nn.init.zeros_(self.fc.bias)
        nn.init.trunc_normal_(self.embedding.weight[:self.num_codes])
        self._pack = None
        self._pack_key = None

    def param_pack(self) -> "native.GnnParamPack":
        """Device-pointer view of the parameters for the C ABI; rebuilt only when a parameter tensor is replaced."""
        if self.identity_proj or self.activation_name != "relu":
            raise NotImplementedError("schemanet_b200 GNN kernels implement the shipped configuration "
                                      "(identity_proj=False, activation='relu'; config/*/schema_net/*.yaml:33-36)")
        tensors = [self.embedding.weight, self.fc.weight, self.fc.bias]
        for layer in self.layers:
            tensors += [layer.g_conv.linear.weight, layer.g_conv.linear.bias, layer.norm.weight, layer.norm.bias]
        key = tuple((t.data_ptr(), t.device) for t in tensors)
        if key != self._pack_key:
            L = self.layers
            self._pack = native.GnnParamPack(
                self.num_codes, self.embed_dim, self.num_layers, self.embedding.weight,
                [l.g_conv.linear.weight for l in L], [l.g_conv.linear.bias for l in L],
                [l.norm.weight for l in L], [l.norm.bias for l in L], self.fc.weight, self.fc.bias, L[0].norm.eps)
            self._pack_key = key
        return self._pack

    def _grad_mode(self, *inputs) -> bool:
        return torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                            any(t is not None and t.requires_grad for t in inputs))

    def _check_inference(self):
        """The packed / fused kernel entry points have no backward; training goes through `forward` (GnnFn)."""
        if self._grad_mode():
            raise NotImplementedError("schemanet_b200: this fused entry point is forward-only; in grad mode call GNN.forward "
                                      "(autograd through schema_inference.graph.autograd.GnnFn) or wrap the call in torch.no_grad()")

    def _forward_kernels(self, nodes, edges, ingredients, feat_mask):
        bs, n = nodes.shape
        sizes = None
        if feat_mask is not None:
            sizes = (n - feat_mask.sum(dim=1)).to(torch.int32)
        edges = edges if edges.is_contiguous() else edges.contiguous()
        nodes = nodes if nodes.is_contiguous() else nodes.contiguous()
        ingredients = ingredients if ingredients.is_contiguous() else ingredients.contiguous()
        return native.gnn_forward(self.param_pack(), bs, n, sizes, ingredients, nodes, n, edges, n * n, n, None,
                                  ws_key=("gnn", n >= 256))

    def forward(self, nodes: torch.Tensor, edges: torch.Tensor, ingredients: torch.LongTensor,
                feat_mask: torch.BoolTensor = None):
        """nodes [bs, n] vertex weights, edges [bs, n, n], ingredients [bs, n] code ids (num_codes = padding),
        feat_mask [bs, n] True at padded TRAILING nodes -> graph embedding [bs, embed_dim].
        In grad mode the result carries autograd history to nodes, edges and every parameter (gnn.py:78-98)."""
        if self._grad_mode(nodes, edges):
            from .autograd import GnnFn
            L = self.layers
            params = [self.embedding.weight, self.fc.weight, self.fc.bias] + [l.g_conv.linear.weight for l in L] + \
                     [l.g_conv.linear.bias for l in L] + [l.norm.weight for l in L] + [l.norm.bias for l in L]
            self.param_pack()                         # (raises for configurations the kernels do not implement)
            return GnnFn.apply(self, feat_mask, ingredients, nodes, edges, *params)
        return self._forward_kernels(nodes.detach(), edges.detach(), ingredients, feat_mask)

    def forward_packed(self, g: "native.PackedGraphs"):
        """Instance graphs straight from stage 2's packed slots: no padding, no host synchronisation; the mean
        divisor (the padded length N = max_b n_b, gnn.py:96) is read on the device.  Forward-only."""
        self._check_inference()
        L = g.L
        return native.gnn_forward(self.param_pack(), g.B, L, g.num_vertices, g.ids, g.vertex_w, L, g.edges, L * L, L,
                                  g.max_vertices, ws_key=("gnn", False))
