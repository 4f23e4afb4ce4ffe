"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors of the unmodified reference, against
the oracle on seeded inputs, and -- at BASELINE.json's full sizes -- through size-independent properties.

Bars (north star): codeword indices, vertex ids and zero patterns bit-exact; vertex weights, edge weights and logits
within 1e-5 relative (fp32 accumulation).  Two metrics, both written here: `rel_close` = max|a-b| / max|ref| per tensor
(logits, class embeddings: entries of one tensor share a scale and cancel against each other); `elem_close` = element by
element, |a-b| <= rtol * max(|ref|, floor * max|ref|) (vertex and edge weights: every entry is a weight in its own
right; the floor, 1e-3 of the tensor's largest entry, only keeps entries that are numerically zero from dividing by ~0).
"""
import os
import sys

import numpy as np
import pytest
import torch

import head_oracle as ho
from conftest import load_golden, split_cat

pytestmark = pytest.mark.gpu
RTOL = 1e-5
HEAD_CASES = ["head_tiny_easy", "head_hard_edge", "head_wide"]


def _t(x, dev="cuda"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def rel_close(a, b, rtol=RTOL, what=""):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol}"


def elem_close(a, b, rtol=RTOL, floor=1e-3, what=""):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    scale = np.maximum(np.abs(b), floor * np.abs(b).max())
    err = (np.abs(a - b) / np.maximum(scale, 1e-300)).max()
    assert err <= rtol, f"{what}: max elementwise relative err = {err:.3e} > {rtol}"


def build_modules(g=None, schema=None, gnn=None, M=None, K=None, Vc=None, D=None, dev="cuda"):
    from schema_inference.graph import SchemaNet, Matcher
    if g is not None:
        _, _, M, K, Vc, D = g["cfg"].tolist()
        schema = {k.split(".", 1)[1]: torch.from_numpy(v) for k, v in g.items() if k.startswith("schema.")}
        gnn = {k.split(".", 1)[1]: torch.from_numpy(v) for k, v in g.items() if k.startswith("gnn.")}
    sn = SchemaNet(M, K, class_max_vertices=Vc, clamp_vertex_attn=-1.0, clamp_edge_attn=-1.0,
                   prune_node_threshold=0.001)
    sn.vertex_weights.copy_(schema["vertex_weights"])
    sn.edge_weights.copy_(schema["edge_weights"])
    sn.vertex_attribute_weights.copy_(schema["w_v"])
    sn.edge_attribute_weights.copy_(schema["w_e"])
    sn.register_class_vertices(schema["class_ingredients"])
    m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=2, identity_proj=False, activation="relu"))
    m.gnn.load_state_dict(gnn)
    return sn.to(dev), m.to(dev)


# ----------------------------------------------------------------------------------------------------------------
# golden vectors of the unmodified reference
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", HEAD_CASES)
def test_golden_discretization_module(name):
    from discretization import Discretization, DiscretizationJitWrapper
    g = load_golden(name)
    M, d = g["vocab"].shape
    disc = Discretization(M, d, uniform_range=[0, 1]).cuda()
    with torch.no_grad():
        disc.vocabulary.weight.copy_(_t(g["vocab"]))
        seq, match = DiscretizationJitWrapper(disc)(_t(g["mid_feat"]))
    assert match.dtype == torch.int64
    assert np.array_equal(match.t().cpu().numpy(), g["ingredients"])          # bit-exact incl. duplicate codewords
    assert np.array_equal(seq.cpu().numpy(), g["seq_out"])


@pytest.mark.parametrize("name", HEAD_CASES)
def test_golden_schema_net_and_matcher_modules(name):
    g = load_golden(name)
    sn, m = build_modules(g)
    attn, attn_cls = _t(g["attn"]), _t(g["attn_cls"])
    with torch.no_grad():
        inst = sn(_t(g["ingredients"]), attn, attn_cls)
    sizes = g["inst_ids_sizes"]
    assert [len(x) for x in inst["instance_ingredients"]] == sizes.tolist()
    assert np.array_equal(torch.cat(inst["instance_ingredients"]).cpu().numpy(), g["inst_ids_cat"])
    rel_close(torch.cat(inst["instance_vertices"]), g["inst_w_cat"], what="vertex weights")
    elem_close(torch.cat(inst["instance_vertices"]), g["inst_w_cat"], what="vertex weights (elementwise)")
    e = torch.cat([x.reshape(-1) for x in inst["instance_edges"]]).cpu().numpy()
    assert np.array_equal(e == 0, g["inst_e_cat"] == 0), "edge zero pattern"
    for a, b in zip(split_cat(e, sizes, True), split_cat(g["inst_e_cat"], sizes, True)):
        rel_close(a, b, what="edges")
        elem_close(a, b, what="edges (elementwise)")
    # the reference's in-place clamp side effect on the caller's tensors (schema_net.py:296,335)
    ref_attn = torch.from_numpy(g["attn"]).clone()
    ref_attn.masked_fill_(ref_attn < -1.0, float("-inf"))
    assert torch.equal(attn.cpu(), ref_attn)
    ref_cls = torch.from_numpy(g["attn_cls"]).clone()
    ref_cls.masked_fill_(ref_cls < -1.0, float("-inf"))
    assert torch.equal(attn_cls.cpu(), ref_cls)
    # atlas + in-place prune of the parameter (schema_net.py:164)
    with torch.no_grad():
        atlas = sn.get_atlas()
    rel_close(atlas["class_vertices"], g["class_vertices"], 2e-6, "class vertices")
    rel_close(atlas["class_edges"], g["class_edges"], 2e-6, "class edges")
    assert np.array_equal(atlas["class_edges"].cpu().numpy() == 0, g["class_edges"] == 0)
    assert np.array_equal(sn.edge_weights.tensor.detach().cpu().numpy(), g["edge_weights_after"])
    # matcher (+ the padded lists it leaves behind, match.py:49-54)
    with torch.no_grad():
        pred = m(inst, atlas)
        f_kg = m.gnn(atlas["class_vertices"], atlas["class_edges"], atlas["class_ingredients"])
    rel_close(f_kg, g["f_kg"], what="class embeddings")
    rel_close(pred, g["pred"], what="logits")
    N = int(g["padded_N"][0])
    assert all(x.shape == (N,) for x in inst["instance_ingredients"])
    assert all(x.shape == (N, N) for x in inst["instance_edges"])
    M = int(g["cfg"][2])
    for i, s in enumerate(sizes.tolist()):
        assert bool((inst["instance_ingredients"][i][s:] == M).all())
        assert float(inst["instance_edges"][i][s:].abs().sum()) == 0.0 and float(inst["instance_edges"][i][:, s:].abs().sum()) == 0.0


@pytest.mark.parametrize("name", HEAD_CASES)
def test_golden_fused_head(name):
    from schemanet_b200.head import SchemaHead
    g = load_golden(name)
    sn, m = build_modules(g)
    head = SchemaHead(_t(g["vocab"]), sn, m)
    out = head(_t(g["mid_feat"]), _t(g["attn"]), _t(g["attn_cls"]))
    assert np.array_equal(out["ingredients"].cpu().numpy(), g["ingredients"])
    rel_close(out["pred"], g["pred"], what="logits (fused head)")
    assert int(out["graphs"].max_vertices) == int(g["inst_ids_sizes"].max())
    # the atlas tensors are only materialised on request; the logits do not depend on it
    assert head.atlas["class_edges"] is None or not head.materialize_atlas
    rel_close(head.atlas["class_vertices"], g["class_vertices"], 2e-6, "class vertices (fused head)")
    head2 = SchemaHead(_t(g["vocab"]), *build_modules(g))
    head2.materialize_atlas = True
    out2 = head2(_t(g["mid_feat"]), _t(g["attn"]), _t(g["attn_cls"]))
    assert torch.equal(out2["pred"], out["pred"])
    rel_close(head2.atlas["class_edges"], g["class_edges"], 2e-6, "class edges (fused head, materialised)")
    assert np.array_equal(head2.atlas["class_edges"].cpu().numpy() == 0, g["class_edges"] == 0)


def test_graphed_head_replays_the_eager_result():
    """GraphedHead (one CUDA graph over both streams) must reproduce the eager head bit for bit, also after the input
    buffers were refilled in place and after a parameter changed between replays."""
    from schemanet_b200.head import SchemaHead, GraphedHead
    g = load_golden("head_tiny_easy")
    sn, m = build_modules(g)
    head = SchemaHead(_t(g["vocab"]), sn, m)
    mid, attn, cls = _t(g["mid_feat"]), _t(g["attn"]), _t(g["attn_cls"])
    eager = head(mid, attn, cls)["pred"].clone()
    rel_close(eager, g["pred"], what="logits (eager, before capture)")
    bufs = (mid.clone(), attn.clone(), cls.clone())
    gh = GraphedHead(head, *bufs)
    assert torch.equal(gh.replay()["pred"], eager)
    # new inputs in the same buffers (a permutation of the batch)
    perm = torch.randperm(mid.shape[1], generator=torch.Generator().manual_seed(3)).cuda()
    bufs[0].copy_(mid[:, perm]); bufs[1].copy_(attn[perm]); bufs[2].copy_(cls[perm])
    assert torch.equal(gh.replay()["pred"], eager[perm])
    # a parameter update between replays is honoured (the graph holds pointers, not values)
    with torch.no_grad():
        m.gnn.fc.bias.add_(0.25)
    want = head(mid[:, perm].contiguous(), attn[perm].contiguous(), cls[perm].contiguous())["pred"].clone()
    assert torch.equal(gh.replay()["pred"], want) and not torch.equal(want, eager[perm])


def test_class_cache_is_keyed_on_parameter_versions():
    """cache_class=True reuses the class embeddings only while the schema / GNN parameters are unchanged (f4)."""
    from schemanet_b200.head import SchemaHead
    from schemanet_b200 import native
    g = load_golden("head_tiny_easy")
    sn, m = build_modules(g)
    head = SchemaHead(_t(g["vocab"]), sn, m)
    args = (_t(g["mid_feat"]), _t(g["attn"]), _t(g["attn_cls"]))
    first = head(*args, cache_class=True)["pred"].clone()
    n0 = native.launch_count()
    again = head(*args, cache_class=True)["pred"].clone()
    cached_launches = native.launch_count() - n0
    assert torch.equal(first, again)
    n0 = native.launch_count()
    head(*args)
    assert native.launch_count() - n0 > cached_launches           # the class side really was skipped
    with torch.no_grad():
        sn.vertex_weights.tensor.mul_(torch.linspace(0.5, 1.5, sn.vertex_weights.tensor.shape[1], device="cuda"))
    changed = head(*args, cache_class=True)["pred"].clone()        # version bump -> recomputed
    assert not torch.equal(changed, first)
    assert torch.equal(changed, head(*args)["pred"])


def test_golden_plain_list_matcher_path():
    """Matcher fed with ordinary Python lists (not SchemaNet's packed output) takes the packing path."""
    g = load_golden("head_tiny_easy")
    sn, m = build_modules(g)
    sizes = g["inst_ids_sizes"]
    inst = {"instance_ingredients": [_t(x) for x in split_cat(g["inst_ids_cat"], sizes)],
            "instance_vertices": [_t(x) for x in split_cat(g["inst_w_cat"], sizes)],
            "instance_edges": [_t(x) for x in split_cat(g["inst_e_cat"], sizes, True)]}
    atlas = {"class_vertices": _t(g["class_vertices"]), "class_edges": _t(g["class_edges"]),
             "class_ingredients": _t(g["schema.class_ingredients"])}
    with torch.no_grad():
        pred = m(inst, atlas)
    rel_close(pred, g["pred"], what="logits (list path)")


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_golden_cpp_extension_dropin(device):
    """The four pybind-compatible functions, with CPU tensors (host entry points) and CUDA tensors."""
    import cpp_extension as ext
    g = load_golden("init_apis")
    B, M, K, Vc = g["cfg"].tolist()
    ing, attn, attn_cls = _t(g["ingredients"], device), _t(g["attn"], device), _t(g["attn_cls"], device)
    geo = _t(g["geo_sim"], device)
    assert np.array_equal(ext.cpp_feat_to_v_attr(ing, attn_cls, M, True).cpu().numpy(), g["v_attr_mean"])
    assert np.array_equal(ext.cpp_feat_to_v_attr(ing, attn_cls, M, False).cpu().numpy(), g["v_attr_sum"])
    assert np.array_equal(ext.cpp_feat_to_v_attr(ing, attn_cls, M, True, True).cpu().numpy(), g["v_attr_only"])
    dicts = [{int(k): v for v, k in enumerate(row)} for row in g["class_ingredients"]]
    label = g["label"].tolist()
    # block sums are accumulated in the reference's order -> bit-exact
    assert np.array_equal(ext.cpp_feat_to_e(ing, attn, geo, dicts, label, Vc, True).cpu().numpy(), g["e_mean"])
    assert np.array_equal(ext.cpp_feat_to_e(ing, attn, geo, dicts, label, Vc, False).cpu().numpy(), g["e_sum"])
    w = torch.tensor([[0.3], [0.7]], device=device)
    cat_ids, cat_w, nv = ext.cpp_feat_to_instance_v(ing, attn_cls, w, False)
    assert nv.device.type == "cpu" and nv.dtype == torch.int64 and cat_ids.device.type == device
    assert np.array_equal(cat_ids.cpu().numpy(), g["iv_sum_ids"]) and np.array_equal(nv.numpy(), g["iv_sum_nv"])
    rel_close(cat_w, g["iv_sum_w"], what="instance_v sum")
    ids = list(torch.split_with_sizes(cat_ids.cpu(), nv.tolist()))
    idicts = [{v: k for k, v in enumerate(i.tolist())} for i in ids]
    es = ext.cpp_feat_to_instance_e(ing, attn, geo, idicts, w, False, False)
    assert all(e.shape == (n, n) for e, n in zip(es, nv.tolist()))
    rel_close(torch.cat([e.reshape(-1) for e in es]), g["ie_sum_cat"], what="instance_e sum")
    # a permuted dictionary is honoured
    perm = [{c: (len(d) - 1 - r) for c, r in d.items()} for d in idicts]
    es_p = ext.cpp_feat_to_instance_e(ing, attn, geo, perm, w, False, False)
    for a, b in zip(es, es_p):
        assert torch.equal(a.flip(0, 1), b)
    # attribute weights that require grad get their gradient through the final mix, like the reference
    wg = torch.tensor([[0.3], [0.7]], device=device, requires_grad=True)
    _, cw, _ = ext.cpp_feat_to_instance_v(ing, attn_cls, wg, True)
    assert cw.requires_grad
    cw.sum().backward()
    assert wg.grad is not None and wg.grad.abs().sum() > 0


def test_instance_graphs_wide_codes_take_the_same_path_result():
    """Codes outside [0, 2^23) (negative, > 2^40) are ranked by the 64-bit sort: an order-preserving relabelling of
    the codes must give bit-identical graphs (only the ids change), for the hot kernel and the init-time APIs."""
    from schemanet_b200 import native
    g = torch.Generator().manual_seed(5)
    B, L, M = 6, 196, 300
    ing = torch.randint(0, M, (B, L), generator=g)
    ing[1] = 7                       # one code everywhere
    ing[2] = torch.arange(L)         # all distinct
    wide = ing * (1 << 33) - (1 << 50)
    attn = (0.5 * torch.randn(B, L, L, generator=g)).cuda()
    cls = (0.5 * torch.randn(B, L, generator=g)).cuda()
    geo = ho.pair_wise_point_sim(14, 14, 1.0, 2.0).cuda()
    w = torch.tensor([0.3, 0.7]).cuda()
    a = native.instance_graphs(ing.cuda(), attn.clone(), cls.clone(), geo, w, w, -1.0, -1.0, zero_pad=True)
    b = native.instance_graphs(wide.cuda(), attn.clone(), cls.clone(), geo, w, w, -1.0, -1.0, zero_pad=True)
    assert torch.equal(a.num_vertices, b.num_vertices)
    nv = a.num_vertices.tolist()
    assert nv[1] == 1 and nv[2] == L
    for i, n in enumerate(nv):
        assert torch.equal(a.ids[i, :n] * (1 << 33) - (1 << 50), b.ids[i, :n])
        assert torch.equal(a.vertex_w[i, :n], b.vertex_w[i, :n])
    assert torch.equal(a.edges, b.edges)
    # and the 64-bit path against the oracle
    wcol = torch.tensor([[0.3], [0.7]])
    eb = b.edges.view(B, L, L)
    ref = ho.instance_graphs(wide, attn.cpu(), cls.cpu(), wcol, wcol, clamp_vertex=-1.0, clamp_edge=-1.0)
    for i, n in enumerate(nv):
        assert torch.equal(b.ids[i, :n].cpu(), ref["instance_ingredients"][i])
        rel_close(b.vertex_w[i, :n], ref["instance_vertices"][i], what="vertices (wide codes)")
        e = eb[i, :n, :n].cpu()
        assert torch.equal(e == 0, ref["instance_edges"][i] == 0)
        rel_close(e, ref["instance_edges"][i], what="edges (wide codes)")
        assert eb[i, n:].abs().sum() == 0 and eb[i, :, n:].abs().sum() == 0


@pytest.mark.parametrize("side,M,B", [(8, 40, 5), (16, 500, 4), (5, 9, 3)])
def test_instance_graphs_other_token_counts(side, M, B):
    """L = 64 / 256 (the 8-columns-per-lane instantiation, L = kMaxL) / 25 tokens against the oracle."""
    from schemanet_b200 import native
    L = side * side
    g = torch.Generator().manual_seed(side)
    ing = torch.randint(0, M, (B, L), generator=g)
    attn = 0.5 * torch.randn(B, L, L, generator=g)
    cls = 0.5 * torch.randn(B, L, generator=g)
    geo = ho.pair_wise_point_sim(side, side, 1.0, 2.0)
    wcol = torch.tensor([[0.4], [0.6]])
    got = native.instance_graphs(ing.cuda(), attn.cuda(), cls.cuda(), geo.cuda(), wcol.cuda(), wcol.cuda(), -0.5, -0.5, zero_pad=True)
    ref = ho.instance_graphs(ing, attn, cls, wcol, wcol, clamp_vertex=-0.5, clamp_edge=-0.5, feat_h=side, feat_w=side)
    e = got.edges.view(B, L, L).cpu()
    for i in range(B):
        n = int(got.num_vertices[i])
        assert n == len(ref["instance_ingredients"][i])
        assert torch.equal(got.ids[i, :n].cpu(), ref["instance_ingredients"][i])
        rel_close(got.vertex_w[i, :n], ref["instance_vertices"][i], what=f"vertices L={L}")
        assert torch.equal(e[i, :n, :n] == 0, ref["instance_edges"][i] == 0)
        rel_close(e[i, :n, :n], ref["instance_edges"][i], what=f"edges L={L}")
        assert e[i, n:].abs().sum() == 0 and e[i, :, n:].abs().sum() == 0


# ----------------------------------------------------------------------------------------------------------------
# seeded inputs vs the oracle
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,d,M,mode", [(8, 192, 128, "easy"), (8, 192, 128, "hard"), (4, 384, 1024, "easy"),
                                        (4, 384, 1024, "hard"), (2, 768, 1000, "hard"), (3, 100, 77, "hard")])
def test_discretize_vs_oracle(B, d, M, mode):
    """Indices must equal torch.cdist(...).argmin on every row that is not fp32-ambiguous; ambiguous rows (fp64
    top-2 gap below 1e-6 relative -- no fp32 summation order can be expected to agree there) are
    counted and must still pick one of the fp64 top-2... and they must be rare."""
    from schemanet_b200 import native
    vocab, mid, _, _ = ho.synth_inputs(B, d, M, seed=1000 + M + d, mode=mode)
    flat = mid[1:].reshape(-1, d)
    want = torch.cdist(flat, vocab).argmin(1)
    got = native.discretize(flat.cuda(), vocab.cuda()).cpu()
    bad = (got != want).nonzero().flatten()
    if len(bad):
        _, gap = ho.discretize_fp64_gap(flat[bad], vocab)
        assert bool((gap < 1e-6).all()), f"{len(bad)} mismatches, some on unambiguous rows (min gap {gap.max():.2e})"
    assert len(bad) <= max(1, flat.shape[0] // 2000)
    if mode == "easy":
        assert len(bad) == 0


@pytest.mark.parametrize("operands", ["tf32", "f16"])
@pytest.mark.parametrize("B,d,M,mode", [(64, 384, 1024, "hard"), (16, 768, 8000, "hard"), (64, 192, 128, "easy"),
                                        (7, 96, 200, "hard")])
def test_discretize_tensor_core_vs_exact_path(B, d, M, mode, operands):
    """tcgen05 (tf32 coarse pass + fp32 re-check) against the fp32 CUDA-core scan on the same device: any
    difference must sit on an fp32-ambiguous row; the re-check statistics are reported."""
    from schemanet_b200 import native
    vocab, mid, _, _ = ho.synth_inputs(B, d, M, seed=2000 + M, mode=mode)
    flat = mid[1:].reshape(-1, d).cuda()
    v = vocab.cuda()
    exact = native.discretize(flat, v, mode=native.DISC_EXACT)
    tc_mode = native.DISC_TENSOR if operands == "tf32" else native.DISC_TENSOR_F16
    tens, ws = native.discretize(flat, v, mode=tc_mode, return_workspace=True)
    stats = native.discretize_stats(ws)
    bad = (exact != tens).nonzero().flatten().cpu()
    print(f"tensor-core ({operands}) discretize B={B} d={d} M={M} {mode}: recheck rows {stats['recheck_rows']} / {flat.shape[0]}, "
          f"overflow rows {stats['overflow_rows']}, mismatches vs exact {len(bad)}")
    if len(bad):
        _, gap = ho.discretize_fp64_gap(flat.cpu()[bad], vocab)
        assert bool((gap < 1e-6).all()), f"{len(bad)} mismatches on unambiguous rows (largest gap {gap.max():.2e})"
    assert len(bad) <= max(1, flat.shape[0] // 2000)
    if operands == "f16":        # (tf32 operands keep 10 mantissa bits: its worst-case band is 8x wider, lists overflow more often)
        assert stats["overflow_rows"] <= flat.shape[0] // 100


@pytest.mark.parametrize("d,M,B,scales,mode", [(384, 1024, 24, [(7, 30.0)], "easy"), (384, 1024, 24, [(7, 30.0)], "hard"),
                                               (384, 1024, 24, [(7, 100.0), (200, 100.0)], "hard"),
                                               (768, 8000, 4, [(5, 30.0)], "hard"), (768, 8000, 4, [(5, 100.0), (300, 60.0)], "easy"),
                                               (192, 128, 16, [(0, 1.0e5)], "hard")])
@pytest.mark.parametrize("operands", ["f16", "tf32"])
def test_discretize_outlier_channels(d, M, B, scales, mode, operands):
    """Non-i.i.d. features: a few channels are 30-100x larger than the rest in the tokens AND in the codebook (DeiT
    layer-9 'massive activations'; the last case exceeds the fp16 range on purpose).  The tensor-core pass only
    short-lists candidates inside a worst-case error band, so every row whose fp64 top-2 gap exceeds 1e-6 must carry
    exactly torch.cdist(...).argmin's index (discretization.py:65) -- ZERO mismatches, no statistical allowance."""
    from schemanet_b200 import native
    vocab, mid, _, _ = ho.synth_inputs(B, d, M, seed=3000 + M + len(scales), mode=mode)
    flat = mid[1:].reshape(-1, d).clone()
    vocab = vocab.clone()
    for ch, s in scales:
        flat[:, ch] *= s
        vocab[:, ch] *= s
    want = torch.cdist(flat, vocab).argmin(1)
    tc_mode = native.DISC_TENSOR_F16 if operands == "f16" else native.DISC_TENSOR
    got, ws = native.discretize(flat.cuda(), vocab.cuda(), mode=tc_mode, return_workspace=True)
    stats = native.discretize_stats(ws)
    exact = native.discretize(flat.cuda(), vocab.cuda(), mode=native.DISC_EXACT).cpu()
    got = got.cpu()
    print(f"outlier channels {scales} ({operands}) d={d} M={M} {mode}: recheck rows {stats['recheck_rows']} / {flat.shape[0]}, "
          f"overflow rows {stats['overflow_rows']}")
    for name, other in (("torch.cdist argmin", want), ("the exact fp32 scan", exact)):
        bad = (got != other).nonzero().flatten()
        if len(bad):                  # two fp32 summation orders may only disagree where fp64 itself sees a tie
            _, gap = ho.discretize_fp64_gap(flat[bad], vocab)
            assert bool((gap < 1e-6).all()), f"{len(bad)} mismatches vs {name} on unambiguous rows (largest gap {gap.max():.2e})"


def test_discretize_cta_pair_variant():
    """SCHEMANET_DISC_CTAS=2 (cta_group::2 pairs, read once per process) must return the same indices: the tensor-core vs
    exact-path cases are re-run in a child process with the variant selected."""
    import subprocess
    env = dict(os.environ, SCHEMANET_DISC_CTAS="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "tensor_core_vs_exact_path or discretize_vs_oracle"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


def test_discretize_edge_cases():
    from schemanet_b200 import native
    g = torch.Generator().manual_seed(5)
    vocab = torch.rand(64, 40, generator=g)
    vocab[40:48] = vocab[8:16]                      # exact duplicates: the lower index must win
    x = torch.cat([vocab, vocab + 1e-4, torch.zeros(3, 40)])
    want = torch.cdist(x, vocab).argmin(1)
    for mode in (native.DISC_AUTO, native.DISC_EXACT, native.DISC_TENSOR, native.DISC_TENSOR_F16):
        got = native.discretize(x.cuda(), vocab.cuda(), mode=mode).cpu()
        assert torch.equal(got, want)
    assert bool((got[40:48] == torch.arange(8, 16)).all())
    # strided output layout used by the head: row r = t*bs + b -> out[b, t]
    bs, L = 5, 7
    x = torch.rand(L * bs, 40, generator=g)
    out = torch.empty(bs, L, dtype=torch.int64, device="cuda")
    native.discretize(x.cuda(), vocab.cuda(), out_idx=out, idx_rows=bs, idx_row_stride=L, idx_col_stride=1)
    assert torch.equal(out.cpu(), torch.cdist(x, vocab).argmin(1).reshape(L, bs).t())
    # a single row, a single codeword
    assert int(native.discretize(torch.rand(1, 40).cuda(), vocab[:1].contiguous().cuda())[0]) == 0


@pytest.mark.parametrize("cfg", ["cfg1"])
def test_full_head_vs_oracle_cfg1(cfg):
    """BASELINE configs[0]: DeiT-Tiny / CIFAR-10 shape, the reference's own CPU-runnable case, end to end."""
    from schemanet_b200.head import SchemaHead
    c = ho.CONFIGS[cfg]
    vocab, mid, attn, attn_cls = ho.synth_inputs(c["B"], c["d"], c["M"], seed=1234)
    schema = ho.synth_schema(c["M"], c["K"], c["Vc"], seed=1235)
    gnn = ho.synth_gnn(c["M"], c["D"], seed=1236)
    ref = ho.head_forward(mid, attn, attn_cls, vocab, schema, gnn, ho.HEAD_CFG)
    sn, m = build_modules(schema=schema, gnn=gnn, M=c["M"], K=c["K"], Vc=c["Vc"], D=c["D"])
    out = SchemaHead(vocab.cuda(), sn, m)(mid.cuda(), attn.cuda(), attn_cls.cuda())
    assert torch.equal(out["ingredients"].cpu(), ref["ingredients"])
    ids, vw, ed, n = out["graphs"].to_lists()
    assert n == [len(x) for x in ref["instance_ingredients"]]
    assert torch.equal(torch.cat(ids).cpu(), torch.cat(ref["instance_ingredients"]))
    rel_close(torch.cat(vw), torch.cat(ref["instance_vertices"]), what="vertices")
    elem_close(torch.cat(vw), torch.cat(ref["instance_vertices"]), what="vertices (elementwise)")
    for a, b in zip(ed, ref["instance_edges"]):
        rel_close(a, b, what="edges")
        elem_close(a, b, what="edges (elementwise)")
    rel_close(out["pred"], ref["pred"], what="logits")
    # the raw-heads entry (stage 0 fused into the graph-build read) gives the same answer
    H = c["H"]
    gen = torch.Generator().manual_seed(99)
    extracted = 0.5 * torch.randn(c["B"] * H, 197, 197, generator=gen)
    a2, c2 = ho.attention_prologue(extracted, c["B"])
    ref2 = ho.head_forward(mid, a2, c2, vocab, schema, gnn, ho.HEAD_CFG)
    out2 = SchemaHead(vocab.cuda(), sn, m)(mid.cuda(), extracted=extracted.cuda())
    rel_close(out2["pred"], ref2["pred"], what="logits (fused prologue)")
    from schemanet_b200 import native
    a3, c3 = native.attention_prologue(extracted.cuda(), c["B"])
    rel_close(a3, a2, 1e-6, "prologue attn")
    rel_close(c3, c2, 1e-6, "prologue attn_cls")


def test_gnn_and_similarity_vs_oracle():
    from schema_inference.graph import Matcher
    gen = torch.Generator().manual_seed(3)
    M, D, bs, n = 50, 96, 5, 37
    params = ho.synth_gnn(M, D, seed=4)
    m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=2)).cuda()
    m.gnn.load_state_dict(params)
    nodes = torch.rand(bs, n, generator=gen)
    edges = torch.rand(bs, n, n, generator=gen)
    ids = torch.randint(0, M, (bs, n), generator=gen)
    sizes = torch.tensor([37, 1, 20, 36, 5])
    mask = torch.arange(n)[None, :] >= sizes[:, None]
    nodes[mask] = 0
    ids[mask] = M
    edges = edges * (~mask)[:, :, None] * (~mask)[:, None, :]
    want = ho.gnn_forward(params, nodes, edges, ids, mask)
    with torch.no_grad():
        got = m.gnn(nodes.cuda(), edges.cuda(), ids.cuda(), mask.cuda())
        got_nomask = m.gnn(nodes.cuda(), edges.cuda(), ids.cuda())
    rel_close(got, want, what="gnn (masked)")
    rel_close(got_nomask, ho.gnn_forward(params, nodes, edges, ids, None), what="gnn (no mask)")
    fk = torch.randn(7, D, generator=gen)
    for kind in ("inner_product", "cosine", "euclidean"):
        mm = Matcher(kind, M, dict(embed_dim=D, num_layers=2)).cuda()
        a, b = want.unsqueeze(1), fk.unsqueeze(0)
        ref = {"inner_product": (a * b).sum(-1), "cosine": (torch.cosine_similarity(a, b, dim=-1) + 1) / 2,
               "euclidean": 1 / (1 + torch.linalg.vector_norm(a - b, dim=-1))}[kind]
        rel_close(mm.similarity(want.cuda(), fk.cuda()), ref, what=kind)


@pytest.mark.parametrize("B,K,D", [(256, 1000, 1024), (2048, 101, 1024), (700, 257, 768)])
def test_similarity_tensor_core_path(B, K, D):
    """ImageNet-scale inner-product logits run as one tensor-core GEMM (three fp16 MMAs per product); K is not a multiple of
    the 256-column tile nor of 4 in two of the cases (ragged output rows)."""
    from schemanet_b200 import native
    gen = torch.Generator().manual_seed(B + K)
    fi = torch.randn(B, D, generator=gen) * 3.0
    fk = torch.randn(K, D, generator=gen) * 0.5 + 0.1
    want = (fi.double() @ fk.double().t()).float()
    got = native.similarity(fi.cuda(), fk.cuda(), "inner_product")
    assert got.shape == (B, K)
    rel_close(got, want, what="logits on the tensor cores")
    # against the one-warp-per-pair kernel on a corner of the problem (same definition, fp32 accumulation)
    small = native.similarity(fi[:8].cuda(), fk[:16].contiguous().cuda(), "inner_product")
    rel_close(got[:8, :16], small, what="tensor-core vs CUDA-core logits")


def test_class_side_large_tile_path():
    """Vc >= 256 takes the 128x128 GEMM tiles; also exercises Vc that is not a multiple of the tile."""
    from schema_inference.graph import Matcher
    gen = torch.Generator().manual_seed(8)
    M, D, K, Vc = 400, 64, 3, 300
    params = ho.synth_gnn(M, D, seed=9)
    m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=2)).cuda()
    m.gnn.load_state_dict(params)
    nodes = torch.rand(K, Vc, generator=gen) / Vc
    edges = torch.rand(K, Vc, Vc, generator=gen) / Vc
    ids = torch.stack([torch.randperm(M, generator=gen)[:Vc] for _ in range(K)])
    with torch.no_grad():
        got = m.gnn(nodes.cuda(), edges.cuda(), ids.cuda())
    rel_close(got, ho.gnn_forward(params, nodes, edges, ids, None), what="gnn class side")


@pytest.mark.parametrize("K,Vc,D,thr,rsl", [(4, 1024, 256, 0.001, False), (3, 500, 256, 0.001, True), (2, 296, 512, 0.0034, False),
                                            (3, 264, 256, None, False), (2, 301, 256, 0.0033, False),
                                            (3, 264, 256, 0.5, False),      # 0.5: every vertex pruned (adjacency = I)
                                            (2, 1056, 256, 0.001, False)])  # Vc > 1024: generic atlas kernel, X0^T by the gather kernel
def test_class_side_fused_equals_atlas_then_gnn(K, Vc, D, thr, rsl):
    """sh_dev_class_side on the tensor-core path feeds the compacted, normalised edges straight into the adjacency
    operand.  It must return bit-identical class embeddings whether or not the full class_edges tensor is asked for,
    the same tensors as sh_dev_class_atlas followed by sh_dev_gnn_forward_class, and leave the same in-place prune."""
    from schemanet_b200 import native
    from schema_inference.graph import GNN
    M = 1200                                         # >= Vc: class ingredients are distinct codes
    sch = ho.synth_schema(M, K, Vc, seed=21)
    gnn = GNN(M, D, num_layers=2).cuda()
    gnn.load_state_dict(ho.synth_gnn(M, D, seed=22))
    vw, ci = sch["vertex_weights"].cuda(), sch["class_ingredients"].cuda()
    ew_a, ew_b, ew_c = (sch["edge_weights"].clone().cuda() for _ in range(3))
    pack = gnn.param_pack()
    cv1, ce1, f1 = native.class_side(pack, vw, ew_a, ci, thr, True, rsl, want_edges=True)
    cv2, ce2, f2 = native.class_side(pack, vw, ew_b, ci, thr, True, rsl, want_edges=False)
    cv3, ce3 = native.class_atlas(vw, ew_c, thr, True, rsl)
    f3 = native.gnn_forward_class(pack, cv3, ce3, ci, thr)
    assert ce2 is None
    assert torch.equal(cv1, cv3) and torch.equal(cv2, cv3) and torch.equal(ce1, ce3)
    assert torch.equal(ew_a, ew_c) and torch.equal(ew_b, ew_c)          # in-place prune of the parameter
    assert torch.equal(f1, f3) and torch.equal(f2, f3)
    # and against the oracle
    atlas = ho.class_atlas(sch["vertex_weights"], sch["edge_weights"].clone(), sch["class_ingredients"], thr, rsl)
    ref = ho.gnn_forward(ho.synth_gnn(M, D, seed=22), atlas["class_vertices"], atlas["class_edges"], sch["class_ingredients"], None)
    rel_close(f2, ref, what="class embeddings (fused class side)")


@pytest.mark.parametrize("K,Vc,masked,D", [(3, 300, False, 256), (5, 1024, False, 256), (9, 196, True, 256), (2, 33, True, 256),
                                          (3, 300, True, 512), (2, 500, False, 1024)])
def test_gnn_tensor_core_path_vs_oracle(K, Vc, masked, D):
    """embed_dim 256 takes the tcgen05 3xTF32 path (adjacency prep, TMA-fed UMMA, LayerNorm fused in the TMEM
    epilogue); it must meet the same 1e-5 bar against the fp32 oracle as the CUDA-core path."""
    from schema_inference.graph import Matcher
    gen = torch.Generator().manual_seed(80 + Vc)
    M = 1500
    params = ho.synth_gnn(M, D, seed=81)
    params["layers.0.norm.weight"] = torch.rand(D, generator=gen) + 0.5
    params["layers.1.norm.bias"] = torch.randn(D, generator=gen) * 0.1
    m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=2)).cuda()
    m.gnn.load_state_dict(params)
    nodes = torch.rand(K, Vc, generator=gen) / Vc
    edges = torch.rand(K, Vc, Vc, generator=gen) / Vc
    ids = torch.stack([torch.randperm(M, generator=gen)[:Vc] for _ in range(K)])
    mask = None
    if masked:
        sizes = torch.randint(1, Vc + 1, (K,), generator=gen)
        sizes[0] = Vc
        mask = torch.arange(Vc)[None, :] >= sizes[:, None]
        nodes[mask] = 0
        ids[mask] = M
        edges = edges * (~mask)[:, :, None] * (~mask)[:, None, :]
    want = ho.gnn_forward(params, nodes, edges, ids, mask)
    with torch.no_grad():
        got = m.gnn(nodes.cuda(), edges.cuda(), ids.cuda(), mask.cuda() if masked else None)
    rel_close(got, want, what="gnn tensor-core path")


@pytest.mark.parametrize("G,n,masked,D,M,layers", [(6, 100, True, 512, 300, 2), (4, 300, False, 1024, 700, 2), (5, 196, True, 1024, 500, 2),
                                                   (3, 160, True, 512, 200, 3), (4, 96, True, 768, 100, 1)])
def test_gnn_wide_fused_path_vs_oracle(G, n, masked, D, M, layers):
    """embed_dim > 256 with at least M + 1 node slots: the first Linear is applied to the embedding table on the tensor
    cores, every GEMM epilogue stores the pre-LayerNorm activations with per-tile statistics, and LayerNorm + ReLU are
    applied by the consumer (the next adjacency GEMM's operand conversion, the pooling).  1, 2 and 3 layers."""
    from schema_inference.graph import Matcher
    gen = torch.Generator().manual_seed(90 + n + D)
    params = ho.synth_gnn(M, D, seed=91, num_layers=layers)
    for i in range(layers):
        params[f"layers.{i}.norm.weight"] = torch.rand(D, generator=gen) + 0.5
        params[f"layers.{i}.norm.bias"] = torch.randn(D, generator=gen) * 0.2
    m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=layers)).cuda()
    m.gnn.load_state_dict(params)
    nodes = torch.rand(G, n, generator=gen) / n
    edges = torch.rand(G, n, n, generator=gen) / n
    ids = torch.randint(0, M, (G, n), generator=gen)
    mask = None
    if masked:
        sizes = torch.randint(1, n + 1, (G,), generator=gen)
        sizes[0] = n
        mask = torch.arange(n)[None, :] >= sizes[:, None]
        nodes[mask] = 0
        ids[mask] = M
        edges = edges * (~mask)[:, :, None] * (~mask)[:, None, :]
    want = ho.gnn_forward(params, nodes, edges, ids, mask, num_layers=layers)
    with torch.no_grad():
        got = m.gnn(nodes.cuda(), edges.cuda(), ids.cuda(), mask.cuda() if masked else None)
    rel_close(got, want, what="gnn wide fused path")


class _FakeBackbone(torch.nn.Module):
    """Stands in for the reference's JIT backbone: returns fixed `mid_feat` / `extracted` taps."""

    def __init__(self, mid, extracted):
        super().__init__()
        self.mid, self.extracted = mid, extracted

    def forward(self, x):
        return {"mid_feat": self.mid, "extracted": self.extracted}


def test_predictor_with_ingredient_wrapper_vs_oracle():
    """SchemaNetPredictor end to end (graph/__init__.py:37-57) on top of IngredientModelWrapper
    (ingredient_model_wrapper.py:43-69): every key of the wrapper's dict and the predictor's dict, requires_graph too."""
    from discretization import Discretization, DiscretizationJitWrapper
    from schema_inference.graph import SchemaNetPredictor
    from schema_inference.utils import IngredientModelWrapper
    B, d, M, K, Vc, D, H = 6, 64, 96, 5, 96, 256, 3
    vocab, mid, _, _ = ho.synth_inputs(B, d, M, seed=501)
    gen = torch.Generator().manual_seed(502)
    extracted = 0.5 * torch.randn(B * H, 197, 197, generator=gen)
    schema = ho.synth_schema(M, K, Vc, seed=503)
    gnn = ho.synth_gnn(M, D, seed=504)
    attn, attn_cls = ho.attention_prologue(extracted, B)
    ref = ho.head_forward(mid, attn, attn_cls, vocab, schema, gnn, ho.HEAD_CFG)
    seq_ref, _ = ho.discretize_with_cls(mid, vocab)

    disc = Discretization(M, d, uniform_range=[0, 1])
    with torch.no_grad():
        disc.vocabulary.weight.copy_(vocab)
    sn, m = build_modules(schema=schema, gnn=gnn, M=M, K=K, Vc=Vc, D=D)
    wrapper = IngredientModelWrapper(_FakeBackbone(mid.cuda(), extracted.cuda()), DiscretizationJitWrapper(disc)).cuda()
    pred = SchemaNetPredictor(wrapper, sn, m).cuda().eval()
    with torch.no_grad():
        w = wrapper(torch.zeros(B, 3, 224, 224, device="cuda"))
        out = pred(torch.zeros(B, 3, 224, 224, device="cuda"), requires_graph=True)
    assert list(w.keys()) == ["cls_token", "feat", "feat_origin", "ingredients", "attn", "attn_cls"]
    assert torch.equal(w["ingredients"].cpu(), ref["ingredients"])
    assert torch.equal(w["cls_token"].cpu(), mid[:1].transpose(0, 1)) and torch.equal(w["feat_origin"].cpu(), mid[1:].transpose(0, 1))
    assert torch.equal(w["feat"].cpu(), seq_ref[1:].transpose(0, 1))
    rel_close(w["attn"], attn, 1e-6, "attn")
    rel_close(w["attn_cls"], attn_cls, 1e-6, "attn_cls")
    assert list(out.keys())[:4] == ["pred", "class_vertices", "class_edges", "class_ingredients"]
    rel_close(out["pred"], ref["pred"], what="predictor logits")
    rel_close(out["class_edges"], ref["class_edges"], 2e-6, "class_edges")
    for k in ("instance_ingredients", "instance_vertices", "instance_edges", "ingredients", "attn_cls"):
        assert k in out


@pytest.mark.parametrize("name,B,K", [("cfg3", 6, 4), ("cfg4", 3, 5), ("cfg4", 3, 64), ("cfg3", 4, 101)])
def test_large_config_slices_vs_oracle(name, B, K):
    """BASELINE configs[2] / configs[3] shapes (DeiT-Base d=768; ImageNet M=8000, Vc=500, D=1024) on a slice of the
    batch and of the class set that the CPU oracle finishes in seconds.  K = 64 at D = 1024 is 64 x 2 row blocks x 4 column
    tiles = 512 work units per class-side launch (wide fused path: every CTA pair walks several tiles, LayerNorm statistics
    merged across the four column tiles); ("cfg3", 4, 101) is the whole Caltech-101 class set."""
    from schemanet_b200.head import SchemaHead
    c = dict(ho.CONFIGS[name], B=B, K=K)
    vocab, mid, attn, attn_cls = ho.synth_inputs(c["B"], c["d"], c["M"], seed=700 + K)
    schema = ho.synth_schema(c["M"], c["K"], c["Vc"], seed=701)
    gnn = ho.synth_gnn(c["M"], c["D"], seed=702)
    ref = ho.head_forward(mid, attn, attn_cls, vocab, schema, gnn, ho.HEAD_CFG)
    sn, m = build_modules(schema=schema, gnn=gnn, M=c["M"], K=c["K"], Vc=c["Vc"], D=c["D"])
    out = SchemaHead(vocab.cuda(), sn, m)(mid.cuda(), attn.cuda(), attn_cls.cuda())
    assert torch.equal(out["ingredients"].cpu(), ref["ingredients"])
    rel_close(out["pred"], ref["pred"], what=f"{name} logits")
    rel_close(out["feat_class"], ho.gnn_forward(gnn, ref["class_vertices"], ref["class_edges"], ref["class_ingredients"]),
              what=f"{name} class embeddings")


def test_full_head_vs_oracle_cfg2():
    """BASELINE configs[1] at FULL size (B=256, d=384, M=1024, K=100, Vc=1024, D=256) -- the configuration bench.py times.
    Every stage against the oracle: 200+ class-side and 256 instance-side GEMM work units over 74 CTA pairs, i.e. the
    persistent multi-tile path (shared-memory ring phases and the TMEM double buffer wrap across tiles)."""
    from schemanet_b200.head import SchemaHead, GraphedHead
    c = ho.CONFIGS["cfg2"]
    vocab, mid, attn, attn_cls = ho.synth_inputs(c["B"], c["d"], c["M"], seed=1234)
    schema = ho.synth_schema(c["M"], c["K"], c["Vc"], seed=1235)
    gnn = ho.synth_gnn(c["M"], c["D"], seed=1236)
    ref = ho.head_forward(mid, attn, attn_cls, vocab, dict(schema, edge_weights=schema["edge_weights"].clone()), gnn, ho.HEAD_CFG)
    sn, m = build_modules(schema=schema, gnn=gnn, M=c["M"], K=c["K"], Vc=c["Vc"], D=c["D"])
    head = SchemaHead(vocab.cuda(), sn, m)
    dev_in = (mid.cuda(), attn.cuda(), attn_cls.cuda())
    out = head(*dev_in)
    assert torch.equal(out["ingredients"].cpu(), ref["ingredients"])
    ids, vw, ed, n = out["graphs"].to_lists()
    assert n == [len(x) for x in ref["instance_ingredients"]]
    assert torch.equal(torch.cat(ids).cpu(), torch.cat(ref["instance_ingredients"]))
    elem_close(torch.cat(vw), torch.cat(ref["instance_vertices"]), what="cfg2 vertices")
    for b in range(0, c["B"], 5):
        assert torch.equal(ed[b].cpu() == 0, ref["instance_edges"][b] == 0), "edge zero pattern"
        elem_close(ed[b], ref["instance_edges"][b], what="cfg2 edges")
    f_kg = ho.gnn_forward(gnn, ref["class_vertices"], ref["class_edges"], ref["class_ingredients"])
    rel_close(out["feat_class"], f_kg, what="cfg2 class embeddings")
    rel_close(out["pred"], ref["pred"], what="cfg2 logits")
    # the CUDA-graph replay bench.py times returns the same logits
    gh = GraphedHead(head, *dev_in)
    rel_close(gh.replay()["pred"], ref["pred"], what="cfg2 logits (graph replay)")


@pytest.mark.parametrize("side,G,n,D,thr", [("class", 48, 1024, 256, 0.001), ("instance", 160, 196, 256, None),
                                            ("class", 24, 500, 1024, 0.001), ("class", 8, 304, 512, 0.0034)])
def test_gnn_more_tiles_than_cta_pairs(side, G, n, D, thr):
    """gemm3x_kernel is persistent: with more work units than CTA pairs (74 on a B200) every CTA walks several tiles.
    class: 48 x 4 = 192 units of 256 rows (fused class side incl. the pruned-vertex tables); instance: 160 graphs;
    wide: D=1024 has 4 column tiles per row block (layer-0 shortcut on the tensor cores, LayerNorm applied by the consumers;
    pruned vertices keep their GEMM rows through the identity tail -- the last case prunes more than half of them)."""
    from schemanet_b200 import native
    from schema_inference.graph import GNN, Matcher
    M = 1200
    params = ho.synth_gnn(M, D, seed=31)
    if side == "class":
        sch = ho.synth_schema(M, G, n, seed=32)
        gnn = GNN(M, D, num_layers=2).cuda()
        gnn.load_state_dict(params)
        _, _, f = native.class_side(gnn.param_pack(), sch["vertex_weights"].cuda(), sch["edge_weights"].clone().cuda(),
                                    sch["class_ingredients"].cuda(), thr, True, False, want_edges=False)
        atlas = ho.class_atlas(sch["vertex_weights"], sch["edge_weights"].clone(), sch["class_ingredients"], thr, False)
        want = ho.gnn_forward(params, atlas["class_vertices"], atlas["class_edges"], sch["class_ingredients"], None)
        rel_close(f, want, what="class side, multi-tile")
    else:
        gen = torch.Generator().manual_seed(33)
        m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=2)).cuda()
        m.gnn.load_state_dict(params)
        sizes = torch.randint(100, n + 1, (G,), generator=gen)
        sizes[0] = n
        mask = torch.arange(n)[None, :] >= sizes[:, None]
        nodes = torch.rand(G, n, generator=gen) / n
        edges = torch.rand(G, n, n, generator=gen) / n
        ids = torch.randint(0, M, (G, n), generator=gen)
        nodes[mask] = 0
        ids[mask] = M
        edges = edges * (~mask)[:, :, None] * (~mask)[:, None, :]
        want = ho.gnn_forward(params, nodes, edges, ids, mask)
        with torch.no_grad():
            got = m.gnn(nodes.cuda(), edges.cuda(), ids.cuda(), mask.cuda())
        rel_close(got, want, what="instance side, multi-tile")


# ----------------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs[1]: B=256, d=384, M=1024) -- no oracle needed
# ----------------------------------------------------------------------------------------------------------------
def test_full_size_properties_cfg2():
    from schemanet_b200 import native
    c = ho.CONFIGS["cfg2"]
    B, d, M = c["B"], c["d"], c["M"]
    gen = torch.Generator(device="cuda").manual_seed(77)
    vocab = torch.rand(M, d, device="cuda", generator=gen)
    pick = torch.randint(0, M, (196 * B,), device="cuda", generator=gen)
    tokens = vocab[pick] + 0.3 * torch.randn(196 * B, d, device="cuda", generator=gen)
    idx = native.discretize(tokens, vocab)
    # (1) optimality: the chosen codeword's distance equals the row minimum of torch.cdist on the device
    rows = torch.randperm(196 * B, device="cuda", generator=gen)[:4096]
    dist = torch.cdist(tokens[rows], vocab)
    chosen = dist.gather(1, idx[rows, None]).squeeze(1)
    assert bool((chosen <= dist.min(1).values * (1 + 1e-6)).all())
    assert float((idx == pick).float().mean()) > 0.999          # easy tokens decode to their generating codeword
    # (2) idempotence: codewords map to themselves
    assert torch.equal(native.discretize(vocab, vocab), torch.arange(M, device="cuda"))
    # (3) graph build: ids strictly ascending and equal to torch.unique; edge rows sum to w0 + w1 (each attribute
    #     channel is row-normalised to 1, large_scale_feat_to_e.cpp:135-140); sizes consistent
    ingredients = idx.reshape(196, B).t().contiguous()
    attn = 0.5 * torch.randn(B, 196, 196, device="cuda", generator=gen)
    attn_cls = 0.5 * torch.randn(B, 196, device="cuda", generator=gen)
    geo = ho.pair_wise_point_sim(14, 14).cuda()
    wv = torch.tensor([0.5, 0.5], device="cuda")
    we = torch.tensor([0.25, 0.75], device="cuda")
    g = native.instance_graphs(ingredients, attn, attn_cls, geo, wv, we, -1.0, -1.0)
    ids, vw, ed, n = g.to_lists()
    assert int(g.max_vertices) == max(n)
    for b in range(0, B, 17):
        assert torch.equal(ids[b], torch.unique(ingredients[b]))
        rel_close(ed[b].sum(1), torch.ones(n[b]), 1e-5, "edge row sums")
        assert float(vw[b].max()) <= 1.0 + 1e-6 and float(vw[b].min()) > 0.0
    # (4) linearity in the attribute weights: e(w) = w0 * e(1,0) + w1 * e(0,1)
    g0 = native.instance_graphs(ingredients, attn, attn_cls, geo, wv, torch.tensor([1.0, 0.0], device="cuda"), -1.0, -1.0)
    g1 = native.instance_graphs(ingredients, attn, attn_cls, geo, wv, torch.tensor([0.0, 1.0], device="cuda"), -1.0, -1.0)
    e0, e1 = g0.to_lists()[2], g1.to_lists()[2]
    for b in range(0, B, 37):
        rel_close(ed[b], 0.25 * e0[b] + 0.75 * e1[b], 1e-6, "linearity")
