"""ctypes binding of libschemahead.so (the C ABI declared in include/schemahead.h).

PyTorch is plumbing here: it owns device memory and streams; every compute call below goes to a hand-written
sm_100a kernel through the C ABI.  There is no CPU fallback: if the library is missing, or there is no CUDA device,
the calls raise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libschemahead.so")
ABI_VERSION = 1

NO_CLAMP = -3.0e38
G_RAW_LOGITS, G_FROM_HEADS, G_SUM, G_WRITE_BACK_CLAMP, G_ZERO_PAD = 1, 2, 4, 8, 16
DISC_AUTO, DISC_EXACT, DISC_TENSOR, DISC_TENSOR_F16 = 0, 1, 2, 3
SIM_KINDS = {"inner_product": 0, "cosine": 1, "euclidean": 2}

# every symbol include/schemahead.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = [
    "sh_abi_version", "sh_last_error", "sh_launch_count", "sh_device_info", "sh_profile_enable", "sh_profile_collect",
    "sh_dev_attention_prologue",
    "sh_discretize_workspace_bytes", "sh_dev_discretize", "sh_discretize_stats",
    "sh_dev_instance_graphs", "sh_dev_feat_to_v_attr", "sh_dev_feat_to_e", "sh_dev_class_accumulate",
    "sh_dev_class_atlas",
    "sh_gnn_workspace_bytes", "sh_dev_gnn_forward", "sh_dev_similarity", "sh_class_side_workspace_bytes",
    "sh_dev_class_side", "sh_dev_gnn_forward_class",
    "sh_host_feat_to_instance_v", "sh_host_feat_to_instance_e", "sh_host_feat_to_v_attr", "sh_host_feat_to_e",
]


class GnnParams(ctypes.Structure):
    _fields_ = [
        ("num_codes", ctypes.c_int), ("embed_dim", ctypes.c_int), ("num_layers", ctypes.c_int),
        ("ln_eps", ctypes.c_float),
        ("embedding", ctypes.c_void_p),
        ("lin_w", ctypes.POINTER(ctypes.c_void_p)), ("lin_b", ctypes.POINTER(ctypes.c_void_p)),
        ("ln_w", ctypes.POINTER(ctypes.c_void_p)), ("ln_b", ctypes.POINTER(ctypes.c_void_p)),
        ("fc_w", ctypes.c_void_p), ("fc_b", ctypes.c_void_p),
    ]


_lib = None


def lib():
    """Load libschemahead.so (fails loudly: the product path has no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        L.sh_last_error.restype = ctypes.c_char_p
        L.sh_launch_count.restype = ctypes.c_int64
        L.sh_discretize_workspace_bytes.restype = ctypes.c_size_t
        L.sh_discretize_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int]
        L.sh_gnn_workspace_bytes.restype = ctypes.c_size_t
        L.sh_gnn_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
        L.sh_dev_attention_prologue.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        L.sh_dev_discretize.argtypes = [vp, vp, i64, i32, i32, vp, i64, i64, i64, vp, vp, ctypes.c_size_t, i32, vp]
        L.sh_discretize_stats.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
        L.sh_dev_instance_graphs.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, f32, vp, vp, i32, vp, vp, vp, vp, vp, vp]
        L.sh_dev_feat_to_v_attr.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
        L.sh_dev_feat_to_e.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
        L.sh_dev_class_accumulate.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp]
        L.sh_dev_class_atlas.argtypes = [vp, vp, i32, i32, f32, i32, i32, vp, vp, vp]
        L.sh_dev_gnn_forward.argtypes = [ctypes.POINTER(GnnParams), i32, i32, vp, vp, vp, i32, vp, i64, i32, vp, vp, vp,
                                         ctypes.c_size_t, vp]
        L.sh_dev_similarity.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
        L.sh_class_side_workspace_bytes.restype = ctypes.c_size_t
        L.sh_class_side_workspace_bytes.argtypes = [i32, i32, i32]
        L.sh_dev_gnn_forward_class.argtypes = [ctypes.POINTER(GnnParams), i32, i32, vp, vp, vp, f32, vp, vp, ctypes.c_size_t, vp]
        L.sh_dev_class_side.argtypes = [ctypes.POINTER(GnnParams), vp, vp, vp, i32, i32, f32, i32, i32, vp, vp, vp, vp,
                                        ctypes.c_size_t, vp]
        L.sh_host_feat_to_instance_v.argtypes = [vp, vp, i32, i32, vp, i32, vp, vp, vp]
        L.sh_host_feat_to_instance_e.argtypes = [vp, vp, vp, i32, i32, vp, i32, vp, vp]
        L.sh_host_feat_to_v_attr.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
        L.sh_host_feat_to_e.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
        if L.sh_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libschemahead ABI {L.sh_abi_version()} != binding {ABI_VERSION}: rebuild")
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libschemahead: " + lib().sh_last_error().decode(errors="replace"))


def launch_count():
    return int(lib().sh_launch_count())


def profile_enable(on=True):
    check(lib().sh_profile_enable(int(on)))


def profile_collect():
    """-> {kernel name: (launches, total milliseconds)} since the last collect; synchronises the device."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().sh_profile_collect(buf, ctypes.c_size_t(len(buf))))
    out = {}
    for line in buf.value.decode().splitlines():
        name, count, ms = line.split("\t")
        out[name] = (int(count), float(ms))
    return out


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("schemanet_b200: this entry point needs CUDA tensors (there is no CPU fallback)")


def _f32c(t):
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected scalar type Float but found {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _i64c(t):
    if t.dtype != torch.int64:   # the reference's `long` accessors raise the same way
        raise RuntimeError(f"expected scalar type Long but found {str(t.dtype).replace('torch.', '').capitalize()}")
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------------------------
# device-level wrappers (CUDA tensors in, CUDA tensors out, asynchronous on the current stream)
# ----------------------------------------------------------------------------------------------------------------
def attention_prologue(extracted, bs):
    """ingredient_model_wrapper.py:57-69: head mean + slicing.  extracted [bs*H, T, T] -> attn [bs,L,L], attn_cls [bs,L]."""
    require_cuda(extracted)
    extracted = _f32c(extracted)
    T = extracted.shape[-1]
    H = extracted.shape[0] // bs
    attn = torch.empty(bs, T - 1, T - 1, device=extracted.device, dtype=torch.float32)
    attn_cls = torch.empty(bs, T - 1, device=extracted.device, dtype=torch.float32)
    check(lib().sh_dev_attention_prologue(ptr(extracted), bs, H, T, ptr(attn), ptr(attn_cls), stream()))
    return attn, attn_cls


_ws_cache = {}


def _workspace(key, nbytes, device):
    """Scratch memory of one kind of call, per device AND per launching stream: two heads driven from two streams of one
    process never share scratch (calls on one stream are ordered, so sharing there is safe)."""
    k = (key, device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(k)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _ws_cache[k] = ws
    return ws


def discretize(tokens, vocab, out_idx=None, idx_rows=None, idx_row_stride=1, idx_col_stride=0, out_seq=None,
               mode=DISC_AUTO, return_workspace=False):
    """discretization.py:65: tokens [R, d], vocab [M, d] -> int64 argmin indices.

    By default returns a flat [R] tensor.  With idx_rows=bs, idx_row_stride=L, idx_col_stride=1 and an `out_idx`
    of shape [bs, L], token-major rows (r = t*bs + b) are written straight into the [bs, L] layout."""
    require_cuda(tokens, vocab)
    tokens, vocab = _f32c(tokens), _f32c(vocab)
    R, d = tokens.shape
    M = vocab.shape[0]
    if vocab.shape[1] != d:
        raise RuntimeError(f"dimension {d} not match to {vocab.shape[1]}")
    if mode == DISC_AUTO and os.environ.get("SCHEMANET_DISC_MODE", "") == "exact":
        mode = DISC_EXACT          # debugging aid: force the fp32 CUDA-core scan
    if out_idx is None:
        out_idx = torch.empty(R, dtype=torch.int64, device=tokens.device)
    if idx_rows is None:
        idx_rows = R
    nbytes = lib().sh_discretize_workspace_bytes(R, d, M)
    ws = _workspace("disc", nbytes, tokens.device)
    check(lib().sh_dev_discretize(ptr(tokens), ptr(vocab), R, d, M, ptr(out_idx), idx_rows, idx_row_stride,
                                  idx_col_stride, ptr(out_seq), ptr(ws), ws.numel(), mode, stream()))
    return (out_idx, ws) if return_workspace else out_idx


def discretize_stats(ws):
    a, b = ctypes.c_int64(0), ctypes.c_int64(0)
    check(lib().sh_discretize_stats(ptr(ws), ctypes.byref(a), ctypes.byref(b)))
    return {"recheck_rows": a.value, "overflow_rows": b.value}


class PackedGraphs:
    """Stage-2 output in the packed slot layout of include/schemahead.h."""
    __slots__ = ("ids", "vertex_w", "edges", "num_vertices", "max_vertices", "B", "L")

    def __init__(self, B, L, device, want_vertices=True, want_edges=True):
        self.B, self.L = B, L
        self.ids = torch.empty(B, L, dtype=torch.int64, device=device) if want_vertices else None
        self.vertex_w = torch.empty(B, L, dtype=torch.float32, device=device) if want_vertices else None
        self.edges = torch.empty(B, L * L, dtype=torch.float32, device=device) if want_edges else None
        self.num_vertices = torch.empty(B, dtype=torch.int32, device=device)
        self.max_vertices = torch.zeros(1, dtype=torch.int32, device=device)

    def to_lists(self):
        """One D2H copy of the B lengths, then zero-copy views (the list API of schema_net.py:302-304)."""
        n = self.num_vertices.tolist()
        ids = [self.ids[b, :n[b]] for b in range(self.B)] if self.ids is not None else None
        vw = [self.vertex_w[b, :n[b]] for b in range(self.B)] if self.vertex_w is not None else None
        L = self.L
        ed = [self.edges[b].view(L, L)[:n[b], :n[b]] for b in range(self.B)] if self.edges is not None else None
        return ids, vw, ed, n


def instance_graphs(ingredients, attn, attn_cls, geo_sim, w_vertex, w_edge, clamp_vertex=None, clamp_edge=None,
                    raw_logits=True, heads=0, mean=True, write_back_clamp=False, zero_pad=False, want_vertices=True,
                    want_edges=True, out=None):
    """schema_net.py:278-375 + large_scale_feat_to_{v,e}.cpp in one launch.  Returns PackedGraphs."""
    require_cuda(ingredients, attn, attn_cls, geo_sim, w_vertex, w_edge)
    ingredients = _i64c(ingredients)
    B, L = ingredients.shape
    flags = (G_RAW_LOGITS if raw_logits else 0) | (G_FROM_HEADS if heads else 0) | (0 if mean else G_SUM) | \
            (G_WRITE_BACK_CLAMP if write_back_clamp else 0) | (G_ZERO_PAD if zero_pad else 0)
    if attn is not None:
        if attn.dtype != torch.float32 or not attn.is_contiguous():
            raise RuntimeError("attn must be a contiguous float32 CUDA tensor")
    if attn_cls is not None:
        if attn_cls.dtype != torch.float32 or not attn_cls.is_contiguous():
            raise RuntimeError("attn_cls must be a contiguous float32 CUDA tensor")
    g = out if out is not None else PackedGraphs(B, L, ingredients.device, want_vertices, want_edges)
    if out is not None:
        g.max_vertices.zero_()
    wv = _f32c(w_vertex.detach().reshape(-1)) if w_vertex is not None else None
    we = _f32c(w_edge.detach().reshape(-1)) if w_edge is not None else None
    geo = _f32c(geo_sim) if geo_sim is not None else None
    check(lib().sh_dev_instance_graphs(
        ptr(ingredients), ptr(attn) if want_edges or heads else None, ptr(attn_cls), ptr(geo) if want_edges else None,
        B, L, heads, NO_CLAMP if clamp_vertex is None else float(clamp_vertex),
        NO_CLAMP if clamp_edge is None else float(clamp_edge), ptr(wv), ptr(we), flags,
        ptr(g.ids), ptr(g.vertex_w), ptr(g.edges) if want_edges else None, ptr(g.num_vertices), ptr(g.max_vertices),
        stream()))
    return g


def feat_to_v_attr(ingredients, attn_cls, n_vertices, mean, ingredients_only):
    require_cuda(ingredients, attn_cls)
    ingredients = _i64c(ingredients)
    B, L = ingredients.shape
    out = torch.empty(B, n_vertices, 2, dtype=torch.float32, device=ingredients.device)
    cls = _f32c(attn_cls) if attn_cls is not None else None
    check(lib().sh_dev_feat_to_v_attr(ptr(ingredients), ptr(cls), B, L, n_vertices, int(mean), int(ingredients_only),
                                      ptr(out), stream()))
    return out


def feat_to_e(ingredients, attn, geo_sim, class_ingredients, label, n_max, mean):
    require_cuda(ingredients, attn, geo_sim, class_ingredients, label)
    ingredients = _i64c(ingredients)
    B, L = ingredients.shape
    K = class_ingredients.shape[0]
    if label.numel() and (int(label.min()) < 0 or int(label.max()) >= K):
        raise IndexError(f"feat_to_e: label out of range [0, {K})")       # the kernel indexes class rows unchecked
    out = torch.empty(B, n_max, n_max, 2, dtype=torch.float32, device=ingredients.device)
    check(lib().sh_dev_feat_to_e(ptr(ingredients), ptr(_f32c(attn)), ptr(_f32c(geo_sim)), ptr(_i64c(class_ingredients)),
                                 ptr(_i64c(label)), B, L, K, n_max, int(mean), ptr(out), stream()))
    return out


def class_accumulate(x, label, acc, n_tracked=None):
    """scripts/init_schema_net.py:31-34,57-59: acc[label[b]] += x[b] (batch order), n_tracked[label[b]] += 1; in place."""
    require_cuda(x, label, acc, n_tracked)
    x, label = _f32c(x), _i64c(label)
    B, K = x.shape[0], acc.shape[0]
    N = x[0].numel()
    if acc.dtype != torch.float32 or not acc.is_contiguous() or acc[0].numel() != N:
        raise RuntimeError("class_accumulate: acc must be a contiguous float32 [K, ...sample shape] tensor")
    if label.numel() != B:
        raise RuntimeError("class_accumulate: one label per sample")
    check(lib().sh_dev_class_accumulate(ptr(x), ptr(label), B, N, K, ptr(acc), ptr(n_tracked), stream()))
    return acc


def class_atlas(vertex_weights, edge_weights, prune_threshold=None, prune_in_place=True, remove_self_loop=False,
                want_edges=True):
    """schema_net.py:144-175.  edge_weights is modified in place where pruned (the reference's side effect, :164)."""
    require_cuda(vertex_weights, edge_weights)
    K, Vc = vertex_weights.shape
    vw = _f32c(vertex_weights.detach())
    ew = edge_weights.detach()
    if ew.dtype != torch.float32 or not ew.is_contiguous():
        raise RuntimeError("edge_weights must be a contiguous float32 CUDA tensor")
    cv = torch.empty(K, Vc, dtype=torch.float32, device=vw.device)
    ce = torch.empty(K, Vc, Vc, dtype=torch.float32, device=vw.device) if want_edges else None
    thr = -1.0 if prune_threshold is None else float(prune_threshold)
    check(lib().sh_dev_class_atlas(ptr(vw), ptr(ew), K, Vc, thr, int(prune_in_place), int(remove_self_loop), ptr(cv),
                                   ptr(ce), stream()))
    return cv, ce


class GnnParamPack:
    """Device pointers of a GNN's parameters in the layout of `sh_gnn_params` (keeps the tensors alive)."""

    def __init__(self, num_codes, embed_dim, num_layers, embedding, lin_w, lin_b, ln_w, ln_b, fc_w, fc_b, ln_eps=1e-5):
        tensors = [embedding, fc_w, fc_b] + list(lin_w) + list(lin_b) + list(ln_w) + list(ln_b)
        require_cuda(*tensors)
        self.keep = [_f32c(t.detach()) for t in tensors]
        emb, fw, fb = self.keep[0], self.keep[1], self.keep[2]
        n = num_layers
        groups = [self.keep[3 + i * n: 3 + (i + 1) * n] for i in range(4)]
        arr_t = ctypes.c_void_p * n
        self.arrays = [arr_t(*[t.data_ptr() for t in g]) for g in groups]
        self.struct = GnnParams(num_codes, embed_dim, num_layers, ln_eps, emb.data_ptr(),
                                ctypes.cast(self.arrays[0], ctypes.POINTER(ctypes.c_void_p)),
                                ctypes.cast(self.arrays[1], ctypes.POINTER(ctypes.c_void_p)),
                                ctypes.cast(self.arrays[2], ctypes.POINTER(ctypes.c_void_p)),
                                ctypes.cast(self.arrays[3], ctypes.POINTER(ctypes.c_void_p)),
                                fw.data_ptr(), fb.data_ptr())
        self.embed_dim = embed_dim
        self.device = emb.device


def gnn_forward(params, G, n_fixed, sizes, ids, vertex_w, ld_v, edges, edge_batch_stride, edge_ld, mean_div, ws_key="gnn"):
    """gnn.py:78-98 for G graphs -> [G, D]."""
    D = params.embed_dim
    out = torch.empty(G, D, dtype=torch.float32, device=params.device)
    nbytes = lib().sh_gnn_workspace_bytes(G, n_fixed, D)
    ws = _workspace(ws_key, nbytes, params.device)
    check(lib().sh_dev_gnn_forward(ctypes.byref(params.struct), G, n_fixed, ptr(sizes), ptr(ids), ptr(vertex_w), ld_v,
                                   ptr(edges), edge_batch_stride, edge_ld, ptr(mean_div), ptr(out), ptr(ws), ws.numel(),
                                   stream()))
    return out


def gnn_tensor_path(embed_dim: int, n_fixed: int) -> bool:
    """Mirror of gnn_tc_supported() (csrc/gnn_tc.cu) plus the debugging overrides: does the GNN run on tcgen05?"""
    return (embed_dim % 256 == 0 and embed_dim <= 1024 and n_fixed >= 32
            and "SCHEMANET_GNN_SIMT" not in os.environ and "SCHEMANET_CLASS_UNFUSED" not in os.environ)


def class_side(params, vertex_weights, edge_weights, class_ingredients, prune_threshold=None, prune_in_place=True,
               remove_self_loop=False, out=None, want_edges=True):
    """get_atlas() + GNN(class graphs) in one call -> (class_vertices, class_edges, feat_class [K, D]).
    want_edges=False (tensor-core path only): the [K, Vc, Vc] class_edges tensor is not materialised (returned as None);
    the normalised edges of the un-pruned vertices go straight into the GNN's adjacency operand."""
    require_cuda(vertex_weights, edge_weights, class_ingredients)
    K, Vc = vertex_weights.shape
    vw = _f32c(vertex_weights.detach())
    ew = edge_weights.detach()
    if ew.dtype != torch.float32 or not ew.is_contiguous():
        raise RuntimeError("edge_weights must be a contiguous float32 CUDA tensor")
    ci = _i64c(class_ingredients)
    if tuple(ci.shape) != (K, Vc) or tuple(ew.shape) != (K, Vc, Vc):
        raise RuntimeError(f"class_side: class_ingredients {tuple(ci.shape)} / edge_weights {tuple(ew.shape)} do not match "
                           f"vertex_weights {(K, Vc)}")
    D = params.embed_dim
    if out is None:
        cv = torch.empty(K, Vc, dtype=torch.float32, device=vw.device)
        ce = torch.empty(K, Vc, Vc, dtype=torch.float32, device=vw.device) if want_edges else None
        fk = torch.empty(K, D, dtype=torch.float32, device=vw.device)
    else:
        cv, ce, fk = out                 # caller-owned buffers (no allocation on this call); ce may be None
    ws = _workspace("class", lib().sh_class_side_workspace_bytes(K, Vc, D), vw.device)
    thr = -1.0 if prune_threshold is None else float(prune_threshold)
    check(lib().sh_dev_class_side(ctypes.byref(params.struct), ptr(vw), ptr(ew), ptr(ci), K, Vc, thr, int(prune_in_place),
                                  int(remove_self_loop), ptr(cv), ptr(ce), ptr(fk), ptr(ws), ws.numel(), stream()))
    return cv, ce, fk


def gnn_forward_class(params, class_vertices, class_edges, class_ingredients, prune_threshold=None):
    """GNN.forward on get_atlas()'s class graphs -> [K, D]; pruned vertices (all-zero edge rows/columns) are skipped."""
    require_cuda(class_vertices, class_edges, class_ingredients)
    K, Vc = class_vertices.shape
    D = params.embed_dim
    out = torch.empty(K, D, dtype=torch.float32, device=class_vertices.device)
    ws = _workspace("class", lib().sh_class_side_workspace_bytes(K, Vc, D), class_vertices.device)
    thr = -1.0 if prune_threshold is None else float(prune_threshold)
    check(lib().sh_dev_gnn_forward_class(ctypes.byref(params.struct), K, Vc, ptr(_f32c(class_vertices.detach())),
                                         ptr(_f32c(class_edges.detach())), ptr(_i64c(class_ingredients)), thr, ptr(out), ptr(ws),
                                         ws.numel(), stream()))
    return out


def similarity(feat_instance, feat_class, kind="inner_product"):
    """match.py:21-31 on the expanded pair -> [B, K]."""
    require_cuda(feat_instance, feat_class)
    B, D = feat_instance.shape
    K = feat_class.shape[0]
    out = torch.empty(B, K, dtype=torch.float32, device=feat_instance.device)
    check(lib().sh_dev_similarity(ptr(_f32c(feat_instance)), ptr(_f32c(feat_class)), B, K, D, SIM_KINDS[kind], ptr(out),
                                  stream()))
    return out


# ----------------------------------------------------------------------------------------------------------------
# host-buffer wrappers (CPU tensors in, CPU tensors out, synchronous): the reference extension's contract
# ----------------------------------------------------------------------------------------------------------------
def host_feat_to_instance_v(ingredients, attn_cls, w_vertex, mean):
    ingredients, attn_cls = _i64c(ingredients), _f32c(attn_cls)
    B, L = ingredients.shape
    ids = torch.empty(B, L, dtype=torch.int64)
    vw = torch.empty(B, L, dtype=torch.float32)
    nv = torch.empty(B, dtype=torch.int64)
    w = _f32c(w_vertex.detach().reshape(-1).cpu())
    check(lib().sh_host_feat_to_instance_v(ptr(ingredients), ptr(attn_cls), B, L, ptr(w), int(mean), ptr(ids), ptr(vw), ptr(nv)))
    return ids, vw, nv


def host_feat_to_instance_e(ingredients, attn, geo_sim, w_edge, mean):
    ingredients, attn, geo_sim = _i64c(ingredients), _f32c(attn), _f32c(geo_sim)
    B, L = ingredients.shape
    e = torch.empty(B, L * L, dtype=torch.float32)
    nv = torch.empty(B, dtype=torch.int64)
    w = _f32c(w_edge.detach().reshape(-1).cpu())
    check(lib().sh_host_feat_to_instance_e(ptr(ingredients), ptr(attn), ptr(geo_sim), B, L, ptr(w), int(mean), ptr(e), ptr(nv)))
    return e, nv


def host_feat_to_v_attr(ingredients, attn_cls, n_vertices, mean, ingredients_only):
    ingredients = _i64c(ingredients)
    B, L = ingredients.shape
    out = torch.empty(B, n_vertices, 2, dtype=torch.float32)
    cls = _f32c(attn_cls) if attn_cls is not None else None
    check(lib().sh_host_feat_to_v_attr(ptr(ingredients), ptr(cls), B, L, n_vertices, int(mean), int(ingredients_only), ptr(out)))
    return out


def host_feat_to_e(ingredients, attn, geo_sim, class_ingredients, label, n_max, mean):
    ingredients, attn, geo_sim = _i64c(ingredients), _f32c(attn), _f32c(geo_sim)
    class_ingredients, label = _i64c(class_ingredients), _i64c(label)
    B, L = ingredients.shape
    K = class_ingredients.shape[0]
    out = torch.empty(B, n_max, n_max, 2, dtype=torch.float32)
    check(lib().sh_host_feat_to_e(ptr(ingredients), ptr(attn), ptr(geo_sim), ptr(class_ingredients), ptr(label), B, L, K,
                                  n_max, int(mean), ptr(out)))
    return out
