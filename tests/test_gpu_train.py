"""GPU: the "next" rows of SURVEY.md section 8 -- f3 (one training iteration: loss terms and every parameter gradient
against the unmodified reference's, fixture tests/golden/train_step.npz), f2 / a11 (atlas initialisation: the init-time
module methods and the per-class accumulation, fixture tests/golden/init_callers.npz) -- plus reference behaviours of the
module path that round 1 did not mirror (remove_self_loop, the expanded similarity call form, label checks)."""
import numpy as np
import pytest
import torch

import head_oracle as ho
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _t(x, dev="cuda"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def rel_close(a, b, rtol, what=""):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if np.isnan(b).any():        # the reference itself emits NaN there (gradients of fully pruned edge rows): same pattern
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
        a, b = np.nan_to_num(a), np.nan_to_num(b)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) if a.size else 0.0
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol}"


def _modules(g, dev="cuda"):
    from schema_inference.graph import SchemaNet, Matcher
    _, _, M, K, Vc, D = g["cfg"].tolist()
    schema = {k.split(".", 1)[1]: torch.from_numpy(v) for k, v in g.items() if k.startswith("schema.")}
    gnn = {k.split(".", 1)[1]: torch.from_numpy(v) for k, v in g.items() if k.startswith("gnn.")}
    sn = SchemaNet(M, K, class_max_vertices=Vc, clamp_vertex_attn=-1.0, clamp_edge_attn=-1.0, prune_node_threshold=0.001)
    sn.vertex_weights.copy_(schema["vertex_weights"]); sn.edge_weights.copy_(schema["edge_weights"])
    sn.vertex_attribute_weights.copy_(schema["w_v"]); sn.edge_attribute_weights.copy_(schema["w_e"])
    sn.register_class_vertices(schema["class_ingredients"])
    m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=2, identity_proj=False, activation="relu"))
    m.gnn.load_state_dict(gnn)
    return sn.to(dev), m.to(dev)


def test_train_step_vs_reference():
    """SchemaNetTrainer.train_iter (tasks/worker_schema_net.py:120-140) on the drop-in modules: normalize(), forward in grad
    mode, SchemaInferenceLoss with the shipped weights, backward.  Logits 1e-5, loss terms 1e-5, gradients 5e-4 (sums of
    cancelling fp32 terms: e.g. d fc.bias = sum of softmax-gradient rows that each add up to zero) of each tensor's largest entry -- against the reference's own autograd."""
    from schema_inference.loss import get_loss_fn
    g = load_golden("train_step")
    sn, m = _modules(g)
    sn.train(); m.train()
    sn.normalize()
    for k, v in sn.state_dict().items():
        rel_close(v, g["after_normalize." + k], 1e-6, "normalize() " + k)
    inst = sn(_t(g["ingredients"]), _t(g["attn"]), _t(g["attn_cls"]))
    atlas = sn.get_atlas()
    pred = m(inst, atlas)
    assert pred.requires_grad
    rel_close(pred, g["pred"], 1e-5, "train-mode logits")
    loss_fn = get_loss_fn({"name": "schema_inference_loss", "loss_cfg": {"re_a_vertex": 3.0, "re_a_edge": 4.0}})
    terms = loss_fn({"pred": pred, **atlas}, {"label": _t(g["label"])})
    for k, v in terms.items():
        ref = float(g["loss." + k][0])
        assert abs(float(v) - ref) <= 1e-5 * max(1.0, abs(ref)), f"loss term {k}: {float(v)} vs {ref}"
    total = terms["cls"] + 0.5 * terms["re_entropy_vertex"] + 0.75 * terms["re_entropy_edge"]
    assert abs(float(total) - float(g["loss_total"][0])) <= 1e-5 * abs(float(g["loss_total"][0]))
    total.backward()
    checked = 0
    for prefix, mod in (("grad.schema_net.", sn), ("grad.matcher.", m)):
        for k, p in mod.named_parameters():
            if p.requires_grad:
                assert p.grad is not None or float(np.abs(g[prefix + k]).max()) == 0.0, f"no gradient for {k}"
                rel_close(p.grad if p.grad is not None else torch.zeros_like(p), g[prefix + k], 5e-4, "grad " + k)
                checked += 1
    assert checked == 4 + 11          # 4 schema parameters; embedding, 2 x (linear w, b, norm w, b), fc w, b
    # pruned COLUMNS of kept rows get exactly zero gradient (schema_net.py:165-166); fully pruned ROWS get NaN in the
    # reference (0/0 behind nan_to_num: a defect kept for parity) unless the module opts out
    keep = (atlas["class_vertices"] > 0.001)
    assert int((~keep).sum()) > 0
    gew = sn.edge_weights.tensor.grad
    assert float(gew[keep.unsqueeze(-1) & ~keep.unsqueeze(-2)].abs().sum()) == 0.0
    assert bool(torch.isnan(gew[~keep]).all())
    ew_after = sn.edge_weights.tensor.detach().cpu().numpy()        # in-place prune of the parameter (schema_net.py:164)
    assert np.array_equal(ew_after == 0, g["after_step.edge_weights.tensor"] == 0)
    rel_close(ew_after, g["after_step.edge_weights.tensor"], 1e-6, "edge_weights after the step")
    sn.zero_grad(); m.zero_grad()
    sn.nan_grad_on_pruned_rows = False
    atlas2 = sn.get_atlas()
    (atlas2["class_edges"] * torch.rand_like(atlas2["class_edges"])).sum().backward()
    assert bool(torch.isfinite(sn.edge_weights.tensor.grad).all()) and float(sn.edge_weights.tensor.grad[~keep].abs().sum()) == 0.0
    sn.nan_grad_on_pruned_rows = True
    sn.zero_grad()
    # an optimiser step afterwards leaves the drop-in usable for inference again
    for p in list(sn.parameters()) + list(m.parameters()):
        if p.grad is None and p.requires_grad:
            p.grad = torch.zeros_like(p)
    torch.optim.SGD([p for p in list(sn.parameters()) + list(m.parameters()) if p.requires_grad], lr=1e-3).step()
    with torch.no_grad():
        sn.normalize()
        out = m(sn(_t(g["ingredients"]), _t(g["attn"]), _t(g["attn_cls"])), sn.get_atlas())
    assert bool(torch.isfinite(out).all())


def test_gnn_autograd_matches_torch_autograd():
    """GnnFn (kernels forward, recompute backward) against plain autograd through the oracle's op chain, masked graphs."""
    from schema_inference.graph import GNN
    gen = torch.Generator().manual_seed(5)
    M, D, bs, n = 60, 256, 4, 40
    params = ho.synth_gnn(M, D, seed=6)
    gnn = GNN(M, D, num_layers=2).cuda()
    gnn.load_state_dict(params)
    sizes = torch.tensor([40, 7, 33, 1])
    mask = torch.arange(n)[None, :] >= sizes[:, None]
    nodes = (torch.rand(bs, n, generator=gen) * ~mask).requires_grad_(True)
    edges = (torch.rand(bs, n, n, generator=gen) * (~mask)[:, :, None] * (~mask)[:, None, :]).requires_grad_(True)
    ids = torch.randint(0, M, (bs, n), generator=gen).masked_fill(mask, M)
    p_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    want = ho.gnn_forward(p_ref, nodes, edges, ids, mask)
    w = torch.randn(bs, D, generator=gen)
    (want * w).sum().backward()
    n2, e2 = nodes.detach().cuda().requires_grad_(True), edges.detach().cuda().requires_grad_(True)
    got = gnn(n2, e2, ids.cuda(), mask.cuda())
    rel_close(got, want, 1e-5, "forward")
    (got * w.cuda()).sum().backward()
    rel_close(n2.grad, nodes.grad, 1e-4, "d nodes")
    rel_close(e2.grad, edges.grad, 1e-4, "d edges")
    for k, p in gnn.named_parameters():
        rel_close(p.grad, p_ref[k].grad if p_ref[k].grad is not None else torch.zeros_like(p_ref[k]), 1e-4, "d " + k)


def _init_batches(g, dev="cuda"):
    nb = int(g["cfg"][0])
    return [{k: _t(g[f"batch{i}.{k}"], dev) for k in ("ingredients", "attn", "attn_cls", "label")} for i in range(nb)]


@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_init_time_module_methods(device):
    """SchemaNet.feat_to_full_vertices / feat_to_limited_edges (schema_net.py:188-274), CUDA tensors (device kernels) and CPU
    tensors (the host-buffer entry points, the reference's own calling convention)."""
    from schema_inference.graph import SchemaNet
    g = load_golden("init_callers")
    nb, B, M, K, Vc = g["cfg"].tolist()
    sn = SchemaNet(M, K, class_max_vertices=Vc, clamp_vertex_attn=-1.0, clamp_edge_attn=-1.0, prune_node_threshold=0.001)
    sn.vertex_attribute_weights.copy_(torch.from_numpy(g["w_v"])); sn.edge_attribute_weights.copy_(torch.from_numpy(g["w_e"]))
    sn = sn.to(device)
    b0 = _init_batches(g, device)[0]
    cls_in = b0["attn_cls"].clone()
    torch.set_grad_enabled(False)                    # (scripts/init_schema_net.py runs under @torch.no_grad())
    v = sn.feat_to_full_vertices(b0["ingredients"], cls_in)
    rel_close(v, g["full_vertices0"], 2e-6, "feat_to_full_vertices")
    assert np.array_equal(v.cpu().numpy() == 0, g["full_vertices0"] == 0)
    ref_cls = torch.from_numpy(g["batch0.attn_cls"]).clone()
    assert torch.equal(cls_in.cpu(), ref_cls.masked_fill_(ref_cls < -1.0, float("-inf")))     # in-place clamp of the caller's tensor (:200-201)
    sn.register_class_vertices(_t(g["valid_vertices"], device))
    e = sn.feat_to_limited_edges(b0["ingredients"], b0["attn"].clone(), b0["label"])
    rel_close(e, g["limited_edges0"], 2e-6, "feat_to_limited_edges")
    assert np.array_equal(e.cpu().numpy() == 0, g["limited_edges0"] == 0)
    with pytest.raises(IndexError):
        sn.feat_to_limited_edges(b0["ingredients"], b0["attn"].clone(), torch.full_like(b0["label"], K))
    torch.set_grad_enabled(True)


def test_atlas_initialisation_passes():
    """scripts/init_schema_net.py:19-65,108-124 on the GPU (schemanet_b200.atlas_init): both dataset passes, the per-class
    running sums by sh_dev_class_accumulate in batch order."""
    from schema_inference.graph import SchemaNet
    from schemanet_b200 import atlas_init, native
    g = load_golden("init_callers")
    nb, B, M, K, Vc = g["cfg"].tolist()
    sn = SchemaNet(M, K, class_max_vertices=Vc, clamp_vertex_attn=-1.0, clamp_edge_attn=-1.0, prune_node_threshold=0.001)
    sn.vertex_attribute_weights.copy_(torch.from_numpy(g["w_v"])); sn.edge_attribute_weights.copy_(torch.from_numpy(g["w_e"]))
    sn.edge_weights.copy_(torch.from_numpy(g["edge_weights_init"]))      # the reference adds to its random initial weights
    sn = sn.cuda()

    class Replay:                                   # re-iterable; hands out copies (the module methods clamp in place)
        def __iter__(self):
            return iter([{k: v.clone() for k, v in b.items()} for b in _init_batches(g)])

    cv = atlas_init.init_class_vertices(Replay(), sn)
    rel_close(cv, g["class_vertices_full"], 2e-6, "init_class_vertices")
    atlas_init.init_schema_net(Replay(), sn)
    assert np.array_equal(sn.class_ingredients.tensor.cpu().numpy(), g["valid_vertices"])
    rel_close(sn.vertex_weights.tensor, g["vertex_weights_final"], 2e-6, "vertex_weights after init")
    rel_close(sn.edge_weights.tensor, g["edge_weights_final"], 2e-6, "edge_weights after init")
    assert sn.class_ingredient_dict[1][int(g["valid_vertices"][1, 3])] == 3
    # the accumulation kernel alone: batch-order sums are bit-exact with a sequential loop
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(9, 5, 7, generator=gen)
    lab = torch.tensor([2, 0, 2, 2, 1, 0, 2, 1, 2])
    acc = torch.rand(3, 5, 7, generator=gen)
    want, cnt = acc.clone(), torch.zeros(3)
    for k, xb in zip(lab.tolist(), x):
        want[k] += xb
        cnt[k] += 1
    acc_d, cnt_d = acc.cuda(), torch.zeros(3, device="cuda")
    native.class_accumulate(x.cuda(), lab.cuda(), acc_d, cnt_d)
    assert torch.equal(acc_d.cpu(), want) and torch.equal(cnt_d.cpu(), cnt)


def test_remove_self_loop_behaves_like_the_reference():
    """remove_self_loop=True: the class atlas drops its diagonal (schema_net.py:170-174); building INSTANCE edges throws in
    the reference (`diagonal(0, 1)`, cpp_extension/src/large_scale_feat_to_e.cpp:136-139) -- and here, on every entry."""
    from schema_inference.graph import SchemaNet, Matcher
    from schemanet_b200.head import SchemaHead
    g = load_golden("head_tiny_easy")
    _, _, M, K, Vc, D = g["cfg"].tolist()
    sn = SchemaNet(M, K, class_max_vertices=Vc, clamp_vertex_attn=-1.0, clamp_edge_attn=-1.0, prune_node_threshold=0.001,
                   remove_self_loop=True).cuda()
    ing, attn, cls = _t(g["ingredients"]), _t(g["attn"]), _t(g["attn_cls"])
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="diagonal dimensions cannot be identical"):
            sn(ing, attn, cls)
        with pytest.raises(RuntimeError, match="diagonal dimensions cannot be identical"):
            sn.feat_to_instance_edges(ing, attn, None)
        ids, vw = sn.feat_to_instance_vertices(ing, cls)          # vertices do not touch the flag
        assert len(ids) == ing.shape[0]
        m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=2)).cuda()
        with pytest.raises(RuntimeError, match="diagonal dimensions cannot be identical"):
            SchemaHead(_t(g["vocab"]), sn, m)(_t(g["mid_feat"]), attn, cls)
        ce = sn.get_atlas()["class_edges"]
        assert float(ce.diagonal(dim1=1, dim2=2).abs().sum()) == 0.0


def test_similarity_accepts_the_reference_call_form():
    """match.py:72-75 calls `self.similarity` on tensors expanded to [bs, K, D]; the compact [bs, D] / [K, D] pair too."""
    from schema_inference.graph import Matcher
    gen = torch.Generator().manual_seed(2)
    a, b = torch.randn(5, 48, generator=gen).cuda(), torch.randn(7, 48, generator=gen).cuda()
    for kind in ("inner_product", "cosine", "euclidean"):
        m = Matcher(kind, 10, dict(embed_dim=48, num_layers=2)).cuda()
        fk = b.expand(5, -1, -1)
        fi = a.unsqueeze(1).expand_as(fk)
        ref = {"inner_product": (fi * fk).sum(-1), "cosine": (torch.cosine_similarity(fi, fk, dim=-1) + 1) / 2,
               "euclidean": 1 / (1 + torch.linalg.vector_norm(fi - fk, dim=-1))}[kind]
        rel_close(m.similarity(fi, fk), ref, 1e-5, kind + " (expanded)")
        rel_close(m.similarity(a, b), ref, 1e-5, kind + " (compact)")
        rel_close(m.similarity(fi.contiguous(), fk.contiguous()), ref, 1e-5, kind + " (materialised)")


def test_host_pipeline_tickets_own_their_slot():
    from schemanet_b200.head import SchemaHead, HostPipeline
    g = load_golden("head_tiny_easy")
    sn, m = _modules(g)
    head = SchemaHead(_t(g["vocab"]), sn, m)
    pipe = HostPipeline(head, torch.device("cuda"), slots=2)
    ins = tuple(torch.from_numpy(g[k]).pin_memory() for k in ("mid_feat", "attn", "attn_cls"))
    t0 = pipe.submit(*ins)
    t1 = pipe.submit(*ins)
    rel_close(pipe.result(t0), g["pred"], 1e-5, "ticket 0")
    t2 = pipe.submit(*ins)                                # reuses ticket 0's slot
    rel_close(pipe.result(t1), g["pred"], 1e-5, "ticket 1")
    rel_close(pipe.result(t2), g["pred"], 1e-5, "ticket 2")
    with pytest.raises(RuntimeError, match="slot was reused"):
        pipe.result(t0)
