"""oracle/head_oracle.py -- CPU restatement of the reference's schema-inference head.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; nothing under schemanet-pytorch_b200/ does (the product path has no CPU fallback).

Every function cites the reference code it restates (paths relative to /root/reference).  Where the reference's
arithmetic *is* a stock ATen CPU op (`torch.cdist`, `softmax`, `bmm`, `linear`, `layer_norm`, `embedding`) the
oracle calls the same ATen op on CPU tensors, because that third-party arithmetic (PyTorch; the reference pins
torch==1.12.1, this image has torch 2.11) is what defines the expected bits.  The reference's native loops
(cpp_extension/src/*.cpp) are restated in plain C in oracle/graph_oracle.c and called through ctypes; when the
reference's own C++ has been compiled into oracle/_ref (oracle/build_ref.py) it can be used instead (`use_ref`).

Parity pinning: the reference has no tests, fixtures or golden vectors.  This oracle is pinned
by (a) tests/golden/*.npz -- outputs of the UNMODIFIED reference classes run in the build container by
oracle/gen_golden.py -- and (b) the reference C++ in oracle/_ref, compared in tests/test_oracle.py.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# --------------------------------------------------------------------------------------------------------------
# native restatement (graph_oracle.c)
# --------------------------------------------------------------------------------------------------------------
def build_c_oracle(force=False):
    """gcc -O2 oracle/graph_oracle.c -> oracle/_build/libgraph_oracle.so"""
    out_dir = os.path.join(HERE, "_build")
    so = os.path.join(out_dir, "libgraph_oracle.so")
    src = os.path.join(HERE, "graph_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        # -ffp-contract=off: no FMA contraction, the reference's scalar fp32 order is kept literally
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def feat_to_instance_v(ingredients, attn_cls, w_v, mean=True):
    """large_scale_feat_to_v.cpp:41-143.  Returns (list ids[n_i] int64, list w[n_i] fp32, num_vertices[B] int64)."""
    ingredients = ingredients.contiguous()
    attn_cls = attn_cls.contiguous().float()
    B, L = ingredients.shape
    ids = torch.zeros(B, L, dtype=torch.int64)
    w = torch.zeros(B, L, dtype=torch.float32)
    nv = torch.zeros(B, dtype=torch.int64)
    wv = w_v.detach().reshape(-1).contiguous().float().cpu()
    _lib().oracle_feat_to_instance_v(_p(ingredients), _p(attn_cls), B, L, _p(wv), int(mean), _p(ids), _p(w), _p(nv))
    n = nv.tolist()
    return [ids[b, :n[b]].clone() for b in range(B)], [w[b, :n[b]].clone() for b in range(B)], nv


def feat_to_instance_e(ingredients, attn, geo_sim, w_e, mean=True):
    """large_scale_feat_to_e.cpp:33-150 with the sorted-rank dictionary of schema_net.py:345-348."""
    ingredients = ingredients.contiguous()
    attn = attn.contiguous().float()
    geo_sim = geo_sim.contiguous().float()
    B, L = ingredients.shape
    e = torch.zeros(B, L * L, dtype=torch.float32)
    nv = torch.zeros(B, dtype=torch.int64)
    we = w_e.detach().reshape(-1).contiguous().float().cpu()
    _lib().oracle_feat_to_instance_e(_p(ingredients), _p(attn), _p(geo_sim), B, L, _p(we), int(mean), _p(e), _p(nv))
    n = nv.tolist()
    return [e[b, :n[b] * n[b]].reshape(n[b], n[b]).clone() for b in range(B)]


def feat_to_v_attr(ingredients, attn_cls, n_vertices, mean=True, ingredients_only=False):
    """feat_to_v_attr.cpp:74-148 -> [B, n_vertices, 2]."""
    ingredients = ingredients.contiguous()
    B, L = ingredients.shape
    attn_cls = attn_cls.contiguous().float() if attn_cls is not None else torch.zeros(B, L)
    out = torch.empty(B, n_vertices, 2, dtype=torch.float32)
    _lib().oracle_feat_to_v_attr(_p(ingredients), _p(attn_cls), B, L, int(n_vertices), int(mean),
                                 int(ingredients_only), _p(out))
    return out


def feat_to_e(ingredients, attn, geo_sim, class_ingredients, label, n_max, mean=True):
    """feat_to_e.cpp:31-127 -> [B, n_max, n_max, 2]."""
    ingredients = ingredients.contiguous()
    attn = attn.contiguous().float()
    geo_sim = geo_sim.contiguous().float()
    class_ingredients = class_ingredients.contiguous()
    label = torch.as_tensor(label, dtype=torch.int64).contiguous()
    B, L = ingredients.shape
    K = class_ingredients.shape[0]
    out = torch.empty(B, n_max, n_max, 2, dtype=torch.float32)
    _lib().oracle_feat_to_e(_p(ingredients), _p(attn), _p(geo_sim), _p(class_ingredients), _p(label),
                            B, L, K, int(n_max), int(mean), _p(out))
    return out


# --------------------------------------------------------------------------------------------------------------
# stage 0: attention prologue  (schema_inference/utils/ingredient_model_wrapper.py:57-69)
# --------------------------------------------------------------------------------------------------------------
def attention_prologue(extracted, bs):
    """extracted [bs*H, T, T] raw logits -> (attn [bs, T-1, T-1], attn_cls [bs, T-1]); head mean then slicing."""
    T = extracted.shape[-1]
    attn = torch.zeros(bs, T, T)
    torch.mean(extracted.unflatten(0, (bs, -1)), dim=1, out=attn)
    return attn[..., 1:, 1:].contiguous(), attn[..., 0, 1:].contiguous()


# --------------------------------------------------------------------------------------------------------------
# stage 1: discretization  (discretization/discretization.py:58-70, visual_word_encoder.py:14-20)
# --------------------------------------------------------------------------------------------------------------
def euclidean_dist_mm(x1, x2):
    """ATen `_euclidean_dist` (the mm path torch.cdist takes for p=2 when P or R > 25), restated op by op:
    sqrt(clamp_min([-2*x1, |x1|^2, 1] @ [x2, 1, |x2|^2]^T, 0)).  On CPU this reproduces torch.cdist bit for bit
    (checked again in tests/test_oracle.py)."""
    x1n = x1.pow(2).sum(-1, keepdim=True)
    x2n = x2.pow(2).sum(-1, keepdim=True)
    a = torch.cat([x1.mul(-2), x1n, torch.ones_like(x1n)], -1)
    b = torch.cat([x2, torch.ones_like(x2n), x2n], -1)
    return a.matmul(b.t()).clamp_min_(0).sqrt_()


def discretize(seq, vocab, activate=True):
    """Discretization.encode (discretization.py:58-70): seq [n, bs, d] -> (seq' [n, bs, d], ingredients [n, bs] int64)."""
    n, bs, d = seq.shape
    flat = seq.detach().reshape(n * bs, d)
    ingredients = torch.cdist(flat, vocab).argmin(dim=1)             # discretization.py:65
    out = F.embedding(ingredients, vocab) if activate else flat      # :66-67
    return out.reshape(n, bs, d), ingredients.reshape(n, bs)


def discretize_with_cls(mid_feat, vocab, activate=True):
    """DiscretizationJitWrapper.forward (scripts/save_backbone_jit.py:127-131): strip the cls token, encode,
    put the cls token back.  mid_feat [1+L, bs, d] -> (seq' [1+L, bs, d], ingredients [L, bs])."""
    seq, ing = discretize(mid_feat[1:], vocab, activate)
    return torch.cat((mid_feat[:1], seq), 0), ing


def discretize_fp64_gap(flat, vocab, chunk=4096):
    """Adjudicator for fp32-ambiguous rows: per row the fp64 argmin of the squared distance and the gap between the two
    smallest fp64 squared distances IN UNITS OF THE REFERENCE FORMULA'S OPERANDS, (d2_2nd - d2_best) / (|x|^2 + |c_best|^2).
    The reference evaluates |x|^2 + |c|^2 - 2 x.c in fp32 (ATen _euclidean_dist: one GEMM over the augmented vectors), so
    what it can resolve is a fraction of |x|^2 + |c|^2, not of the distance itself: on features with large common
    components (outlier channels) two codewords whose distances differ by 70 % can still be indistinguishable to it.
    Rows whose gap is ~1e-6 or less cannot be matched bit for bit by ANY fp32 implementation with a different summation
    order (hard part 1); torch.cdist itself disagrees with fp64 on such rows."""
    v = vocab.double()
    vn = (v * v).sum(-1)
    idx, gap = [], []
    for s in range(0, flat.shape[0], chunk):
        x = flat[s:s + chunk].double()
        xn = (x * x).sum(-1, keepdim=True)
        d2 = xn + vn[None, :] - 2.0 * x @ v.t()
        top = torch.topk(d2, 2, dim=1, largest=False)
        idx.append(top.indices[:, 0])
        gap.append((top.values[:, 1] - top.values[:, 0]) / (xn[:, 0] + vn[top.indices[:, 0]]).clamp_min(1e-300))
    return torch.cat(idx), torch.cat(gap)


# --------------------------------------------------------------------------------------------------------------
# stage 2: instance graphs  (schema_inference/graph/schema_net.py:278-399, graph/utils.py:55-81)
# --------------------------------------------------------------------------------------------------------------
def pair_wise_point_sim(h, w, alpha=1.0, pow=2.0):
    """graph/utils.py:55-81: G[p,q] = 1 / (1 + cdist(grid)[p,q] / alpha)."""
    i, j = torch.meshgrid(torch.arange(h, dtype=torch.float), torch.arange(w, dtype=torch.float), indexing="ij")
    p = torch.stack((i.flatten(), j.flatten()), dim=1)
    return 1 / (1 + torch.cdist(p, p, p=pow) / alpha)


def instance_graphs(ingredients, attn, attn_cls, w_v, w_e, clamp_vertex=None, clamp_edge=None,
                    feat_h=14, feat_w=14, dist_alpha=1.0, dist_pow=2.0, ext=None):
    """SchemaNet.forward (schema_net.py:377-399).  `attn` / `attn_cls` are raw logits and are NOT modified
    (the reference masked_fill_s its caller's tensors in place, :296,:335 -- the oracle works on copies).
    ext: None -> C restatement; or the reference's own compiled `extension` module (oracle/_ref)."""
    attn_cls = attn_cls.clone()
    attn = attn.clone()
    if clamp_vertex is not None:
        attn_cls.masked_fill_(attn_cls < clamp_vertex, float("-inf"))        # :295-296
    a_cls = attn_cls.softmax(dim=-1).nan_to_num(0)                           # :297
    if clamp_edge is not None:
        attn.masked_fill_(attn < clamp_edge, float("-inf"))                  # :334-335
    a = torch.softmax(attn, dim=-1)                                          # :336
    geo = pair_wise_point_sim(feat_h, feat_w, dist_alpha, dist_pow)          # :337-343
    if ext is None:
        ids, wv, nv = feat_to_instance_v(ingredients, a_cls, w_v, mean=True)
        edges = feat_to_instance_e(ingredients, a, geo, w_e, mean=True)
    else:
        cat_ids, cat_w, nv = ext.feat_to_instance_v(ingredients, a_cls, w_v.detach(), True)
        sizes = nv.tolist()
        ids = list(torch.split_with_sizes(cat_ids, sizes))
        wv = list(torch.split_with_sizes(cat_w, sizes))
        dicts = [{v: k for k, v in enumerate(i.tolist())} for i in ids]      # :345-348
        edges = ext.feat_to_instance_e(ingredients, a, geo, dicts, w_e.detach(), True, False)
    return {"instance_ingredients": ids, "instance_vertices": wv, "instance_edges": edges}


# --------------------------------------------------------------------------------------------------------------
# stage 3a: class atlas  (schema_net.py:144-184, graph/utils.py:25-52)
# --------------------------------------------------------------------------------------------------------------
def class_atlas(vertex_weights, edge_weights, class_ingredients, prune_node_threshold=None, remove_self_loop=False):
    """SchemaNet.get_atlas.  Returns the dict and (like the reference, :164) ALSO zeroes the pruned entries of
    `edge_weights` in place -- pass a clone to keep the caller's tensor."""
    def normalize_sum_clamp(x, min_val=0.0):
        x = x.clamp_min(min_val)
        return (x / x.sum(dim=-1, keepdim=True)).nan_to_num(0)
    cv = normalize_sum_clamp(vertex_weights, 1.0e-5)                                     # :144-150
    ew = edge_weights
    if prune_node_threshold is not None:
        mask = (cv > prune_node_threshold).float().unsqueeze(-1)                         # :157-163
        mask = torch.bmm(mask, mask.transpose(1, 2))
        ew.masked_fill_(~mask.bool(), 0)                                                 # :164 (in place)
        ew = ew * mask                                                                   # :166
    ce = normalize_sum_clamp(ew)                                                         # :168
    if remove_self_loop:
        m = torch.ones_like(ce)
        m.diagonal(dim1=1, dim2=2).fill_(0)
        ce = ce * m
    return {"class_vertices": cv, "class_edges": ce, "class_ingredients": class_ingredients}


# --------------------------------------------------------------------------------------------------------------
# stage 3b: GNN + matcher  (schema_inference/graph/gnn.py:7-98, match.py:33-76)
# --------------------------------------------------------------------------------------------------------------
def gnn_forward(params, nodes, edges, ingredients, feat_mask=None, num_layers=2, eps=1e-5):
    """GNN.forward (gnn.py:78-98) with relu activation, identity_proj=False.  params: state-dict style keys
    embedding.weight, layers.{i}.g_conv.linear.{weight,bias}, layers.{i}.norm.{weight,bias}, fc.{weight,bias}."""
    feat = F.embedding(ingredients, params["embedding.weight"])                           # :91
    D = feat.shape[-1]
    for i in range(num_layers):
        adj = edges + edges.transpose(1, 2)                                               # gnn.py:27
        eye = torch.zeros_like(adj)
        eye.diagonal(dim1=1, dim2=2).fill_(1)
        feat = torch.bmm(adj / 2 + eye, feat)                                             # :30
        feat = F.linear(feat, params[f"layers.{i}.g_conv.linear.weight"], params[f"layers.{i}.g_conv.linear.bias"])
        if feat_mask is not None:
            feat = feat.masked_fill(feat_mask[..., None], 0)                              # :43-44
        feat = F.relu(F.layer_norm(feat, (D,), params[f"layers.{i}.norm.weight"], params[f"layers.{i}.norm.bias"], eps))
    feat = (feat * nodes[..., None]).mean(dim=1)                                          # :95-96 (mean over PADDED n)
    return F.linear(feat, params["fc.weight"], params["fc.bias"])                         # :97


def pad_instance_graphs(instance_dict, num_codes):
    """Matcher.forward's padding (match.py:41-54), without mutating the caller's lists."""
    ids, wv, ed = instance_dict["instance_ingredients"], instance_dict["instance_vertices"], instance_dict["instance_edges"]
    sizes = [len(x) for x in ids]
    N = max(sizes)
    bs = len(ids)
    mask = torch.zeros(bs, N, dtype=torch.bool)
    pid = torch.full((bs, N), num_codes, dtype=torch.int64)
    pw = torch.zeros(bs, N)
    pe = torch.zeros(bs, N, N)
    for i, s in enumerate(sizes):
        mask[i, s:] = True
        pid[i, :s] = ids[i]
        pw[i, :s] = wv[i]
        pe[i, :s, :s] = ed[i]
    return pid, pw, pe, mask


def match(gnn_params, instance_dict, class_dict, num_codes, num_layers=2, similarity="inner_product"):
    """Matcher.forward (match.py:33-76) -> logits [bs, K]."""
    pid, pw, pe, mask = pad_instance_graphs(instance_dict, num_codes)
    f_inst = gnn_forward(gnn_params, pw, pe, pid, mask, num_layers)                                  # :56-61
    f_kg = gnn_forward(gnn_params, class_dict["class_vertices"], class_dict["class_edges"],
                       class_dict["class_ingredients"], None, num_layers)                             # :66-70
    a = f_inst.unsqueeze(1)
    b = f_kg.unsqueeze(0)
    if similarity == "inner_product":
        return (a * b).sum(-1)                                                                        # :29-31
    if similarity == "cosine":
        return (torch.cosine_similarity(a, b, dim=-1) + 1) / 2                                        # :21-23
    if similarity == "euclidean":
        return 1 / (1 + torch.linalg.vector_norm(a - b, dim=-1))                                      # :25-27
    raise KeyError(similarity)


def head_forward(mid_feat, attn, attn_cls, vocab, schema, gnn_params, cfg, ext=None):
    """The whole head after the backbone (graph/__init__.py:37-57 minus the JIT backbone call).
    schema: dict(vertex_weights, edge_weights, class_ingredients, w_v, w_e); cfg: dict of SchemaNet/GNN options."""
    _, ing = discretize_with_cls(mid_feat, vocab)
    ingredients = ing.t().contiguous()                                         # ingredient_model_wrapper.py:55
    inst = instance_graphs(ingredients, attn, attn_cls, schema["w_v"], schema["w_e"],
                           cfg.get("clamp_vertex_attn"), cfg.get("clamp_edge_attn"), ext=ext)
    atlas = class_atlas(schema["vertex_weights"], schema["edge_weights"].clone(), schema["class_ingredients"],
                        cfg.get("prune_node_threshold"), cfg.get("remove_self_loop", False))
    logits = match(gnn_params, inst, atlas, vocab.shape[0], cfg.get("num_layers", 2))
    return {"pred": logits, "ingredients": ingredients, **inst, **atlas}


# --------------------------------------------------------------------------------------------------------------
# atlas initialisation  (schema_net.py:188-274, scripts/init_schema_net.py:19-65)
# --------------------------------------------------------------------------------------------------------------
def feat_to_full_vertices(ingredients, attn_cls, num_vertices, w_v, clamp_vertex=None):
    """SchemaNet.feat_to_full_vertices (schema_net.py:188-207) -> [bs, M].  Works on a copy of attn_cls."""
    attn_cls = attn_cls.clone()
    if clamp_vertex is not None:
        attn_cls.masked_fill_(attn_cls < clamp_vertex, float("-inf"))
    attrs = feat_to_v_attr(ingredients, attn_cls.softmax(dim=-1), num_vertices, mean=True)          # :202
    attrs = (attrs / attrs.max(dim=1, keepdim=True)[0]).nan_to_num(0)                                # normalize_max_, :204
    return (attrs @ w_v.reshape(2, 1)).squeeze(-1)


def feat_to_limited_edges(ingredients, attn, class_ingredients, label, w_e, clamp_edge=None, remove_self_loop=False,
                          feat_h=14, feat_w=14, dist_alpha=1.0, dist_pow=2.0):
    """SchemaNet.feat_to_limited_edges (schema_net.py:222-254) -> [bs, Vc, Vc].  Works on a copy of attn."""
    attn = attn.clone()
    if clamp_edge is not None:
        attn.masked_fill_(attn < clamp_edge, float("-inf"))
    geo = pair_wise_point_sim(feat_h, feat_w, dist_alpha, dist_pow)
    attrs = feat_to_e(ingredients, torch.softmax(attn, dim=-1), geo, class_ingredients, label, class_ingredients.shape[1], True)
    attrs = (attrs / attrs.sum(dim=2, keepdim=True)).nan_to_num(0)                                   # normalize_sum_, :248
    if remove_self_loop:
        attrs.diagonal(dim1=1, dim2=2).fill_(0)
    return (attrs @ w_e.reshape(2, 1)).squeeze(-1)


def init_class_vertices(batches, num_classes, num_vertices, w_v, clamp_vertex=None):
    """scripts/init_schema_net.py:43-65: per-class mean of the full vertex weights, rows normalised to sum 1."""
    acc = torch.zeros(num_classes, num_vertices)
    n_tracked = torch.zeros(num_classes)
    for b in batches:
        v = feat_to_full_vertices(b["ingredients"], b["attn_cls"], num_vertices, w_v, clamp_vertex)
        for cls_id, inst in zip(b["label"].tolist(), v):                                             # :57-59, in order
            acc[cls_id] += inst
            n_tracked[cls_id] += 1
    acc /= n_tracked[:, None]
    acc /= acc.sum(dim=-1, keepdim=True)
    return acc


def init_graph(batches, edge_weights, class_ingredients, w_e, clamp_edge=None, remove_self_loop=False):
    """scripts/init_schema_net.py:19-40 up to (not including) graph.normalize(): edge_weights (modified in place) += the
    class-local edges of every sample, then / n_tracked."""
    n_tracked = torch.zeros(edge_weights.shape[0])
    for b in batches:
        e = feat_to_limited_edges(b["ingredients"], b["attn"], class_ingredients, b["label"], w_e, clamp_edge, remove_self_loop)
        for cls_id, inst in zip(b["label"].tolist(), e):                                             # :32-34
            edge_weights[cls_id] += inst
            n_tracked[cls_id] += 1
    edge_weights /= n_tracked[:, None, None]
    return edge_weights


def schema_normalize(schema, apply_normalize=True, clamp_weights=True, remove_self_loop=False):
    """SchemaNet.normalize (schema_net.py:131-142), in place on the dict's tensors."""
    with torch.no_grad():
        if clamp_weights:
            schema["w_v"].clamp_(min=0.01, max=10)
            schema["w_e"].clamp_(min=0.01, max=10)
        if apply_normalize:
            for k in ("vertex_weights", "edge_weights"):
                x = schema[k].clamp_min_(0)
                x /= x.sum(dim=-1, keepdim=True)
                x.nan_to_num_(0)
            if remove_self_loop:
                schema["edge_weights"].diagonal(dim1=1, dim2=2).fill_(0)
    return schema


# --------------------------------------------------------------------------------------------------------------
# training step  (tasks/worker_schema_net.py:120-140, loss/schema_inference_loss.py:21-67)
# --------------------------------------------------------------------------------------------------------------
def class_atlas_train(vertex_weights, edge_weights, class_ingredients, prune_node_threshold=None, remove_self_loop=False):
    """get_atlas in grad mode: the row sums are DETACHED (normalize_sum_clamp(detach_sum=True), schema_net.py:149,168), the
    pruned entries are zeroed in place under no_grad and masked again so that their gradient is zero (:157-166)."""
    def nsc(x, min_val=0.0):
        x = x.clamp_min(min_val)
        return (x / x.sum(dim=-1, keepdim=True).detach()).nan_to_num(0)
    cv = nsc(vertex_weights, 1.0e-5)
    ew = edge_weights
    if prune_node_threshold is not None:
        with torch.no_grad():
            mask = (nsc(vertex_weights.detach(), 1.0e-5) > prune_node_threshold).float().unsqueeze(-1)
            mask = torch.bmm(mask, mask.transpose(1, 2))
            edge_weights.masked_fill_(~mask.bool(), 0)
        ew = edge_weights * mask
    ce = nsc(ew)
    if remove_self_loop:
        m = torch.ones_like(ce)
        m.diagonal(dim1=1, dim2=2).fill_(0)
        ce = ce * m
    return {"class_vertices": cv, "class_edges": ce, "class_ingredients": class_ingredients}


def entropy(p, eps=1.0e-7):
    return -torch.sum(p * torch.log(p + eps), dim=-1)                                                # loss :50-57


def rectify_linear(x, a=0.0):
    return x if x > a else a - 1 + 1.0 / (1 + a - x)                                                  # loss :60-67


def schema_inference_loss(pred, class_vertices, class_edges, label, re_a_vertex=3.0, re_a_edge=3.0):
    """SchemaInferenceLoss.forward (loss/schema_inference_loss.py:21-47)."""
    ev = entropy(class_vertices).max(dim=0)[0]
    ee = entropy(class_edges).max(dim=1)[0].mean()
    return {"cls": F.cross_entropy(pred, label), "entropy_vertex": ev, "entropy_edge": ee,
            "re_entropy_vertex": rectify_linear(ev, re_a_vertex), "re_entropy_edge": rectify_linear(ee, re_a_edge)}


def train_forward(ingredients, attn, attn_cls, schema, gnn_params, cfg):
    """The head's forward in grad mode on leaf tensors that require grad (schema: vertex_weights, edge_weights, w_v, w_e;
    gnn_params).  The 2 -> 1 attribute mixes stay in torch so that autograd reaches w_v / w_e, like the reference's trailing
    matmul (large_scale_feat_to_v.cpp:124-125, large_scale_feat_to_e.cpp:135-140)."""
    e0, e1 = torch.tensor([1.0, 0.0]), torch.tensor([0.0, 1.0])
    g0 = instance_graphs(ingredients, attn, attn_cls, e0, e0, cfg.get("clamp_vertex_attn"), cfg.get("clamp_edge_attn"))
    g1 = instance_graphs(ingredients, attn, attn_cls, e1, e1, cfg.get("clamp_vertex_attn"), cfg.get("clamp_edge_attn"))
    wv, we = schema["w_v"].reshape(-1), schema["w_e"].reshape(-1)
    inst = {"instance_ingredients": g0["instance_ingredients"],
            "instance_vertices": [a * wv[0] + b * wv[1] for a, b in zip(g0["instance_vertices"], g1["instance_vertices"])],
            "instance_edges": [a * we[0] + b * we[1] for a, b in zip(g0["instance_edges"], g1["instance_edges"])]}
    atlas = class_atlas_train(schema["vertex_weights"], schema["edge_weights"], schema["class_ingredients"],
                              cfg.get("prune_node_threshold"), cfg.get("remove_self_loop", False))
    # match() pads with plain tensor assignment, which keeps the graph
    pid, pw, pe, mask = pad_instance_graphs(inst, gnn_params["embedding.weight"].shape[0] - 1)
    f_inst = gnn_forward(gnn_params, pw, pe, pid, mask, cfg.get("num_layers", 2))
    f_kg = gnn_forward(gnn_params, atlas["class_vertices"], atlas["class_edges"], atlas["class_ingredients"], None,
                       cfg.get("num_layers", 2))
    pred = (f_inst.unsqueeze(1) * f_kg.unsqueeze(0)).sum(-1)
    return pred, atlas


# --------------------------------------------------------------------------------------------------------------
# seeded synthetic inputs -- shared by gen_golden.py, the tests and bench.py
# --------------------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: B, d, H, M, K, Vc, D
    "cfg1": dict(B=64, d=192, H=3, M=128, K=10, Vc=128, D=256),
    "cfg2": dict(B=256, d=384, H=6, M=1024, K=100, Vc=1024, D=256),
    "cfg3": dict(B=512, d=768, H=12, M=1024, K=101, Vc=1024, D=256),
    "cfg4": dict(B=1024, d=768, H=12, M=8000, K=1000, Vc=500, D=1024),
}
HEAD_CFG = dict(clamp_vertex_attn=-1.0, clamp_edge_attn=-1.0, prune_node_threshold=0.001,
                remove_self_loop=False, num_layers=2)


def synth_inputs(B, d, M, seed, L=196, mode="easy", vocab=None):
    """vocab ~ U[0,1); tokens: easy = codeword + 0.3 N(0,1), hard = U[0,1); attention logits ~ 0.5 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    if vocab is None:
        vocab = torch.rand(M, d, generator=g)
    if mode == "easy":
        pick = torch.randint(0, M, (L + 1, B), generator=g)
        mid = vocab[pick] + 0.3 * torch.randn(L + 1, B, d, generator=g)
    else:
        mid = torch.rand(L + 1, B, d, generator=g)
    attn = 0.5 * torch.randn(B, L, L, generator=g)
    attn_cls = 0.5 * torch.randn(B, L, generator=g)
    return vocab, mid.contiguous(), attn, attn_cls


def synth_schema(M, K, Vc, seed):
    """Seeded stand-in for SchemaNet._reset_parameters (schema_net.py:104-119) + register_class_vertices."""
    g = torch.Generator().manual_seed(seed)
    vw = torch.empty(K, Vc).normal_(0.5, 1 / 6, generator=g).clamp_(0, 1)
    ew = torch.empty(K, Vc, Vc).normal_(0.5, 1 / 6, generator=g).clamp_(0, 1)
    vw = (vw / vw.sum(-1, keepdim=True)).nan_to_num(0)
    ew = (ew / ew.sum(-1, keepdim=True)).nan_to_num(0)
    ci = torch.stack([torch.randperm(M, generator=g)[:Vc] for _ in range(K)])
    return dict(vertex_weights=vw, edge_weights=ew, class_ingredients=ci,
                w_v=torch.full((2, 1), 0.5), w_e=torch.full((2, 1), 0.5))


def synth_gnn(M, D, seed, num_layers=2):
    """Seeded stand-in for GNN._reset_parameters (gnn.py:73-76) and GraphConv._reset_parameters (:16-19)."""
    g = torch.Generator().manual_seed(seed)
    p = {"embedding.weight": torch.zeros(M + 1, D)}
    p["embedding.weight"][:M] = torch.empty(M, D).normal_(0, 1, generator=g).clamp_(-2, 2)
    bound = (6.0 / (D + D)) ** 0.5
    for i in range(num_layers):
        p[f"layers.{i}.g_conv.linear.weight"] = (torch.rand(D, D, generator=g) * 2 - 1) * bound
        p[f"layers.{i}.g_conv.linear.bias"] = torch.randn(D, generator=g)
        p[f"layers.{i}.norm.weight"] = torch.ones(D)
        p[f"layers.{i}.norm.bias"] = torch.zeros(D)
    p["fc.weight"] = torch.randn(D, D, generator=g)
    p["fc.bias"] = torch.zeros(D)
    return p
