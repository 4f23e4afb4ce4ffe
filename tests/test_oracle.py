"""CPU: pin the oracle (oracle/head_oracle.py + oracle/graph_oracle.c) against the golden vectors that the
UNMODIFIED reference produced (oracle/gen_golden.py), and against the reference's own C++ in oracle/_ref."""
import numpy as np
import pytest
import torch

import head_oracle as ho
from conftest import load_golden, split_cat

HEAD_CASES = ["head_tiny_easy", "head_hard_edge", "head_wide"]
RTOL = 1e-5   # north star: weights/logits within 1e-5 relative


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def _schema(g):
    return {k.split(".", 1)[1]: _t(v) for k, v in g.items() if k.startswith("schema.")}


def _gnn(g):
    return {k.split(".", 1)[1]: _t(v) for k, v in g.items() if k.startswith("gnn.")}


def rel_close(a, b, rtol=RTOL, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    if np.isnan(b).any():        # the reference itself emits NaN there (e.g. gradients of fully pruned edge rows): same pattern
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
        a, b = np.nan_to_num(a), np.nan_to_num(b)
    scale = np.maximum(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale if a.size else 0.0
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e}"


@pytest.mark.parametrize("name", HEAD_CASES)
def test_discretize_matches_reference(name):
    g = load_golden(name)
    seq, ing = ho.discretize_with_cls(_t(g["mid_feat"]), _t(g["vocab"]))
    assert np.array_equal(ing.t().numpy(), g["ingredients"])          # bit-exact indices
    assert np.array_equal(seq.numpy(), g["seq_out"])


def test_euclidean_mm_restatement_is_cdist():
    g = load_golden("head_wide")
    x = _t(g["mid_feat"])[1:].reshape(-1, g["mid_feat"].shape[-1])
    v = _t(g["vocab"])
    assert torch.equal(ho.euclidean_dist_mm(x, v), torch.cdist(x, v))


@pytest.mark.parametrize("name", HEAD_CASES)
def test_instance_graphs_match_reference(name):
    g = load_golden(name)
    sc = _schema(g)
    inst = ho.instance_graphs(_t(g["ingredients"]), _t(g["attn"]), _t(g["attn_cls"]), sc["w_v"], sc["w_e"],
                              ho.HEAD_CFG["clamp_vertex_attn"], ho.HEAD_CFG["clamp_edge_attn"])
    sizes = g["inst_ids_sizes"]
    assert [len(x) for x in inst["instance_ingredients"]] == sizes.tolist()
    assert np.array_equal(torch.cat(inst["instance_ingredients"]).numpy(), g["inst_ids_cat"])
    rel_close(torch.cat(inst["instance_vertices"]).numpy(), g["inst_w_cat"], what="vertex weights")
    e = torch.cat([x.reshape(-1) for x in inst["instance_edges"]]).numpy()
    assert np.array_equal(e == 0, g["inst_e_cat"] == 0)              # zero pattern bit-exact
    for a, b in zip(split_cat(e, sizes, True), split_cat(g["inst_e_cat"], sizes, True)):
        rel_close(a, b, what="edges")


@pytest.mark.parametrize("name", HEAD_CASES)
def test_atlas_and_match_match_reference(name):
    g = load_golden(name)
    sc = _schema(g)
    ew = sc["edge_weights"].clone()
    atlas = ho.class_atlas(sc["vertex_weights"], ew, sc["class_ingredients"], ho.HEAD_CFG["prune_node_threshold"])
    assert np.array_equal(atlas["class_vertices"].numpy(), g["class_vertices"])
    assert np.array_equal(atlas["class_edges"].numpy(), g["class_edges"])
    assert np.array_equal(ew.numpy(), g["edge_weights_after"])       # in-place prune side effect (schema_net.py:164)
    sizes = g["inst_ids_sizes"]
    inst = {"instance_ingredients": [_t(x) for x in split_cat(g["inst_ids_cat"], sizes)],
            "instance_vertices": [_t(x) for x in split_cat(g["inst_w_cat"], sizes)],
            "instance_edges": [_t(x) for x in split_cat(g["inst_e_cat"], sizes, True)]}
    M = int(g["cfg"][2])
    pred = ho.match(_gnn(g), inst, atlas, M)
    rel_close(pred.numpy(), g["pred"], what="logits")


@pytest.mark.parametrize("name", HEAD_CASES)
def test_head_end_to_end(name):
    g = load_golden(name)
    out = ho.head_forward(_t(g["mid_feat"]), _t(g["attn"]), _t(g["attn_cls"]), _t(g["vocab"]), _schema(g), _gnn(g),
                          ho.HEAD_CFG)
    rel_close(out["pred"].numpy(), g["pred"], what="logits e2e")


def test_init_time_apis():
    g = load_golden("init_apis")
    ing, attn, attn_cls = _t(g["ingredients"]), _t(g["attn"]), _t(g["attn_cls"])
    B, M, K, Vc = g["cfg"].tolist()
    assert np.array_equal(ho.feat_to_v_attr(ing, attn_cls, M, True).numpy(), g["v_attr_mean"])
    assert np.array_equal(ho.feat_to_v_attr(ing, attn_cls, M, False).numpy(), g["v_attr_sum"])
    assert np.array_equal(ho.feat_to_v_attr(ing, attn_cls, M, True, True).numpy(), g["v_attr_only"])
    geo = _t(g["geo_sim"])
    ci, label = _t(g["class_ingredients"]), _t(g["label"])
    assert np.array_equal(ho.feat_to_e(ing, attn, geo, ci, label, Vc, True).numpy(), g["e_mean"])
    assert np.array_equal(ho.feat_to_e(ing, attn, geo, ci, label, Vc, False).numpy(), g["e_sum"])
    w = torch.tensor([[0.3], [0.7]])
    ids, wv, nv = ho.feat_to_instance_v(ing, attn_cls, w, mean=False)
    assert np.array_equal(torch.cat(ids).numpy(), g["iv_sum_ids"])
    assert np.array_equal(nv.numpy(), g["iv_sum_nv"])
    rel_close(torch.cat(wv).numpy(), g["iv_sum_w"], what="iv sum")
    es = ho.feat_to_instance_e(ing, attn, geo, w, mean=False)
    rel_close(torch.cat([e.reshape(-1) for e in es]).numpy(), g["ie_sum_cat"], what="ie sum")


def test_geo_sim_closed_form():
    g = load_golden("init_apis")
    assert np.array_equal(ho.pair_wise_point_sim(14, 14, 1, 2).numpy(), g["geo_sim"])
    # closed form used by the CUDA fast path: 1 / (1 + sqrt(dy^2 + dx^2))
    p = np.arange(196)
    dy = (p[:, None] // 14 - p[None, :] // 14).astype(np.float32)
    dx = (p[:, None] % 14 - p[None, :] % 14).astype(np.float32)
    closed = (np.float32(1) / (np.float32(1) + np.sqrt(dy * dy + dx * dx, dtype=np.float32))).astype(np.float32)
    assert np.array_equal(closed, g["geo_sim"])


def test_c_restatement_vs_reference_cpp():
    """oracle/_ref = the reference's own C++; the C restatement must agree with it on fresh random inputs."""
    import build_ref
    ext = build_ref.load()
    if ext is None:
        pytest.skip("oracle/_ref not built (only buildable where /root/reference exists)")
    gen = torch.Generator().manual_seed(7)
    B, L, M = 5, 196, 300
    ing = torch.randint(0, M, (B, L), generator=gen)
    a = torch.softmax(torch.randn(B, L, L, generator=gen), -1)
    ac = torch.softmax(torch.randn(B, L, generator=gen), -1)
    w = torch.tensor([[0.4], [1.1]])
    geo = ho.pair_wise_point_sim(14, 14)
    cat_ids, cat_w, nv = ext.feat_to_instance_v(ing, ac, w, True)
    ids, wv, nv2 = ho.feat_to_instance_v(ing, ac, w, True)
    assert torch.equal(nv, nv2) and torch.equal(cat_ids, torch.cat(ids))
    rel_close(torch.cat(wv).numpy(), cat_w.numpy(), rtol=2e-7, what="v")
    dicts = [{v: k for k, v in enumerate(i.tolist())} for i in ids]
    es_ref = ext.feat_to_instance_e(ing, a, geo, dicts, w, True, False)
    es = ho.feat_to_instance_e(ing, a, geo, w, True)
    for x, y in zip(es, es_ref):
        rel_close(x.numpy(), y.numpy(), rtol=5e-7, what="e")


def _init_batches(g):
    nb = int(g["cfg"][0])
    return [{k: _t(g[f"batch{i}.{k}"]) for k in ("ingredients", "attn", "attn_cls", "label")} for i in range(nb)]


def test_init_callers_match_reference():
    """scripts/init_schema_net.py's two dataset passes over SchemaNet.feat_to_full_vertices / feat_to_limited_edges
    (fixture: the reference's own functions and classes, oracle/gen_golden.py::run_init_callers_case)."""
    g = load_golden("init_callers")
    nb, B, M, K, Vc = g["cfg"].tolist()
    w_v, w_e = _t(g["w_v"]), _t(g["w_e"])
    batches = _init_batches(g)
    clamp = ho.HEAD_CFG["clamp_vertex_attn"]
    full0 = ho.feat_to_full_vertices(batches[0]["ingredients"], batches[0]["attn_cls"], M, w_v, clamp)
    rel_close(full0.numpy(), g["full_vertices0"], 2e-6, "feat_to_full_vertices")
    cv = ho.init_class_vertices(batches, K, M, w_v, clamp)
    rel_close(cv.numpy(), g["class_vertices_full"], 2e-6, "init_class_vertices")
    init_w, valid = cv.topk(Vc, dim=1)
    assert np.array_equal(valid.numpy(), g["valid_vertices"])
    ci = _t(g["valid_vertices"])
    lim0 = ho.feat_to_limited_edges(batches[0]["ingredients"], batches[0]["attn"], ci, batches[0]["label"], w_e,
                                    ho.HEAD_CFG["clamp_edge_attn"])
    rel_close(lim0.numpy(), g["limited_edges0"], 2e-6, "feat_to_limited_edges")
    assert np.array_equal(lim0.numpy() == 0, g["limited_edges0"] == 0)
    schema = dict(vertex_weights=_t(g["vertex_weights"]).clone(), edge_weights=_t(g["edge_weights_init"]).clone(),
                  w_v=w_v.clone(), w_e=w_e.clone())
    ho.init_graph(batches, schema["edge_weights"], ci, w_e, ho.HEAD_CFG["clamp_edge_attn"])
    ho.schema_normalize(schema)
    rel_close(schema["edge_weights"].numpy(), g["edge_weights_final"], 2e-6, "init_graph + normalize: edges")
    rel_close(schema["vertex_weights"].numpy(), g["vertex_weights_final"], 2e-6, "init_graph + normalize: vertices")


def test_train_step_matches_reference():
    """One training iteration (normalize, forward in grad mode, SchemaInferenceLoss, backward) against the loss terms and
    parameter gradients of the unmodified reference classes (oracle/gen_golden.py::run_train_case)."""
    g = load_golden("train_step")
    schema, gnn = _schema(g), _gnn(g)
    schema = {k: v.clone() for k, v in schema.items()}
    ho.schema_normalize(schema)
    names = {"vertex_weights": "vertex_weights.tensor", "edge_weights": "edge_weights.tensor",
             "w_v": "vertex_attribute_weights.tensor", "w_e": "edge_attribute_weights.tensor"}
    for k, ref_name in names.items():
        assert np.array_equal(schema[k].numpy(), g["after_normalize." + ref_name])
        schema[k].requires_grad_(True)
    gnn = {k: v.clone().requires_grad_(True) for k, v in gnn.items()}
    pred, atlas = ho.train_forward(_t(g["ingredients"]), _t(g["attn"]), _t(g["attn_cls"]), schema, gnn, ho.HEAD_CFG)
    rel_close(pred.detach().numpy(), g["pred"], what="train-mode logits")
    terms = ho.schema_inference_loss(pred, atlas["class_vertices"], atlas["class_edges"], _t(g["label"]), 3.0, 4.0)
    for k, v in terms.items():
        assert abs(v.item() - g["loss." + k][0]) <= 1e-5 * max(1.0, abs(g["loss." + k][0])), k
    total = terms["cls"] + 0.5 * terms["re_entropy_vertex"] + 0.75 * terms["re_entropy_edge"]
    total.backward()
    # gradient bar: 1e-4 of the tensor's largest entry (gradients are sums of cancelling fp32 terms -- the two attribute
    # weights in particular, whose channels are each row-normalised; the forward quantities above keep the 1e-5 bar)
    for k, ref_name in names.items():
        rel_close(schema[k].grad.numpy(), g["grad.schema_net." + ref_name], 1e-4, "grad " + k)
    for k, v in gnn.items():
        rel_close((v.grad if v.grad is not None else torch.zeros_like(v)).numpy(), g["grad.matcher.gnn." + k], 1e-4, "grad gnn." + k)
