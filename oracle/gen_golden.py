#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference head in the build container.

The reference has no golden vectors of its own; these fixtures are outputs of its real classes
(`Discretization`, `SchemaNet`, `Matcher`, and the four `cpp_extension` functions compiled from its own C++ by
oracle/build_ref.py) on small seeded inputs.  The fixtures travel to the GPU box; this script and /root/reference
do not need to.

Usage: python oracle/gen_golden.py          (writes tests/golden/*.npz; deterministic)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import head_oracle as ho      # noqa: E402  (only for the seeded input generators)
import ref_harness            # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
L = 196


def load_params(ns, M, d, K, Vc, D, vocab, schema, gnn):
    disc = ns.Discretization(M, d, uniform_range=[0, 1])
    with torch.no_grad():
        disc.vocabulary.weight.copy_(vocab)
    sn = ns.SchemaNet(M, K, class_max_vertices=Vc, **{k: v for k, v in ho.HEAD_CFG.items() if k != "num_layers"})
    sn.vertex_weights.copy_(schema["vertex_weights"])
    sn.edge_weights.copy_(schema["edge_weights"])
    sn.vertex_attribute_weights.copy_(schema["w_v"])
    sn.edge_attribute_weights.copy_(schema["w_e"])
    sn.register_class_vertices(schema["class_ingredients"])
    m = ns.Matcher("inner_product", M, dict(embed_dim=D, num_layers=2, identity_proj=False, activation="relu"))
    m.gnn.load_state_dict(gnn)
    return disc, sn, m


def pack_lists(prefix, lst, out):
    out[prefix + "_sizes"] = np.array([int(x.shape[0]) for x in lst], dtype=np.int64)
    out[prefix + "_cat"] = torch.cat([x.detach().reshape(-1) for x in lst]).numpy()


def run_head_case(ns, name, B, d, M, K, Vc, D, seed, mode, edit=None, w_v=None, w_e=None):
    vocab, mid, attn, attn_cls = ho.synth_inputs(B, d, M, seed, L, mode)
    schema = ho.synth_schema(M, K, Vc, seed + 1)
    if w_v is not None:
        schema["w_v"] = torch.tensor(w_v).reshape(2, 1)
    if w_e is not None:
        schema["w_e"] = torch.tensor(w_e).reshape(2, 1)
    gnn = ho.synth_gnn(M, D, seed + 2)
    if edit is not None:
        vocab, mid, attn, attn_cls = edit(vocab, mid, attn, attn_cls)
    disc, sn, matcher = load_params(ns, M, d, K, Vc, D, vocab, schema, gnn)
    out = dict(vocab=vocab.numpy(), mid_feat=mid.numpy(), attn=attn.numpy(), attn_cls=attn_cls.numpy(),
               cfg=np.array([B, d, M, K, Vc, D], dtype=np.int64))
    for k, v in schema.items():
        out["schema." + k] = v.numpy()
    for k, v in gnn.items():
        out["gnn." + k] = v.numpy()
    with torch.no_grad():
        # stage 1 exactly as DiscretizationJitWrapper.forward (scripts/save_backbone_jit.py:127-131)
        ad = ns.Adapter()
        seq, match = disc(ad.adapt(mid))
        seq, match = ad.reconstruct(seq, match)
        ingredients = match.t().contiguous()
        out["seq_out"] = seq.numpy()
        out["ingredients"] = ingredients.numpy()
        # stage 2 (the reference clamps its inputs in place -> hand it copies)
        inst = sn(ingredients, attn.clone(), attn_cls.clone())
        pack_lists("inst_ids", inst["instance_ingredients"], out)
        pack_lists("inst_w", inst["instance_vertices"], out)
        out["inst_e_cat"] = torch.cat([e.reshape(-1) for e in inst["instance_edges"]]).numpy()
        # stage 3a
        ew_before = sn.edge_weights.tensor.detach().clone()
        atlas = sn.get_atlas()
        out["class_vertices"] = atlas["class_vertices"].numpy()
        out["class_edges"] = atlas["class_edges"].numpy()
        out["edge_weights_after"] = sn.edge_weights.tensor.detach().numpy().copy()
        out["pruned_entries"] = np.array([(ew_before != sn.edge_weights.tensor).sum().item()], dtype=np.int64)
        # stage 3b (Matcher pads the instance lists in place -> they were packed above)
        pred = matcher(inst, atlas)
        out["pred"] = pred.numpy()
        N = inst["instance_ingredients"][0].shape[0]
        out["padded_N"] = np.array([N], dtype=np.int64)
        # intermediate embeddings, for finer-grained parity checks
        out["f_kg"] = matcher.gnn(atlas["class_vertices"], atlas["class_edges"], atlas["class_ingredients"]).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    sizes = out["inst_ids_sizes"]
    print(f"{name}: B={B} d={d} M={M} K={K} Vc={Vc} D={D}  n_i min/max={sizes.min()}/{sizes.max()} "
          f"pruned={out['pruned_entries'][0]} pred[0,:3]={out['pred'][0, :3]}")


def edit_edge_cases(vocab, mid, attn, attn_cls):
    """Edge cases of SURVEY.md section 8d: duplicated codewords (exact ties -> lowest index), a single-code image,
    an all-distinct image (n = 196), a fully masked attention row and a fully masked cls-attention vector."""
    M = vocab.shape[0]
    vocab = vocab.clone()
    vocab[M // 2:M // 2 + 8] = vocab[3:11]                         # exact duplicates: indices 3..10 must win
    mid = mid.clone()
    mid[1:, 0] = vocab[5][None, :]                                 # image 0: every token sits on codeword 5
    mid[1:, 1] = vocab[torch.arange(20, 20 + L)]                   # image 1: 196 distinct codewords
    mid[1:, 2] = vocab[M // 2 + 2][None, :] + 1e-3                 # image 2: near a duplicated pair (3+2 vs M/2+2)
    attn = attn.clone()
    attn[1, 7, :] = -3.0                                           # below clamp everywhere -> softmax NaN row
    attn[2, :, 11] = -2.0                                          # a fully masked column (exact zeros)
    attn_cls = attn_cls.clone()
    attn_cls[3, :] = -5.0                                          # fully masked -> NaN -> nan_to_num(0)
    return vocab, mid, attn, attn_cls


def run_init_case(ns, name, B, M, K, Vc, seed):
    """Init-time dense APIs (schema_net.py:188-274; feat_to_v_attr.cpp, feat_to_e.cpp)."""
    g = torch.Generator().manual_seed(seed)
    ing = torch.randint(0, M, (B, L), generator=g)
    attn = torch.softmax(0.5 * torch.randn(B, L, L, generator=g), -1)
    attn_cls = torch.softmax(0.5 * torch.randn(B, L, generator=g), -1)
    label = torch.randint(0, K, (B,), generator=g)
    ci = torch.stack([torch.randperm(M, generator=g)[:Vc] for _ in range(K)])
    geo = ns.graph_utils.pair_wise_point_sim(14, 14, 1, 2)
    dicts = [{k.item(): v for v, k in enumerate(row)} for row in ci]
    ext = ns.cpp_extension
    out = dict(ingredients=ing.numpy(), attn=attn.numpy(), attn_cls=attn_cls.numpy(), label=label.numpy(),
               class_ingredients=ci.numpy(), geo_sim=geo.numpy(), cfg=np.array([B, M, K, Vc], dtype=np.int64))
    out["v_attr_mean"] = ext.cpp_feat_to_v_attr(ing, attn_cls, M, True, False).numpy()
    out["v_attr_sum"] = ext.cpp_feat_to_v_attr(ing, attn_cls, M, False, False).numpy()
    out["v_attr_only"] = ext.cpp_feat_to_v_attr(ing, attn_cls, M, True, True).numpy()
    out["e_mean"] = ext.cpp_feat_to_e(ing, attn, geo, dicts, label.tolist(), Vc, True).numpy()
    out["e_sum"] = ext.cpp_feat_to_e(ing, attn, geo, dicts, label.tolist(), Vc, False).numpy()
    # instance-level functions with mean=False as well (the Python callers always pass True)
    w = torch.tensor([[0.3], [0.7]])
    cat_ids, cat_w, nv = ext.cpp_feat_to_instance_v(ing, attn_cls, w, False)
    out["iv_sum_ids"], out["iv_sum_w"], out["iv_sum_nv"] = cat_ids.numpy(), cat_w.numpy(), nv.numpy()
    ids = list(torch.split_with_sizes(cat_ids, nv.tolist()))
    idicts = [{v: k for k, v in enumerate(i.tolist())} for i in ids]
    es = ext.cpp_feat_to_instance_e(ing, attn, geo, idicts, w, False, False)
    out["ie_sum_cat"] = torch.cat([e.reshape(-1) for e in es]).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: B={B} M={M} K={K} Vc={Vc}")


def run_init_callers_case(ns, name, nb, B, M, K, Vc, seed):
    """Atlas initialisation as scripts/init_schema_net.py:19-65,108-124 runs it: the reference's own init_class_vertices /
    init_graph (imported unmodified; the DataLoader and the backbone wrapper are stand-ins that replay fixed batches)
    over the reference SchemaNet.feat_to_full_vertices / feat_to_limited_edges (schema_net.py:188-274)."""
    import importlib
    init_mod = importlib.import_module("scripts.init_schema_net")
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)                      # SchemaNet._reset_parameters draws the initial edge_weights
    sn = ns.SchemaNet(M, K, class_max_vertices=Vc, **{k: v for k, v in ho.HEAD_CFG.items() if k != "num_layers"})
    sn.vertex_attribute_weights.copy_(torch.tensor([[0.25], [1.5]]))
    sn.edge_attribute_weights.copy_(torch.tensor([[2.0], [0.125]]))
    batches = []
    for i in range(nb):
        label = torch.randint(0, K, (B,), generator=g)
        label[:K] = torch.randperm(K, generator=g)           # every class is seen (n_tracked > 0)
        batches.append(dict(ingredients=torch.randint(0, M, (B, L), generator=g),
                            attn=0.5 * torch.randn(B, L, L, generator=g), attn_cls=0.5 * torch.randn(B, L, generator=g),
                            label=label))
    out = dict(cfg=np.array([nb, B, M, K, Vc], dtype=np.int64), w_v=sn.vertex_attribute_weights.tensor.detach().numpy().copy(),
               w_e=sn.edge_attribute_weights.tensor.detach().numpy().copy())
    for i, b in enumerate(batches):
        for k, v in b.items():
            out[f"batch{i}.{k}"] = v.numpy().copy()
    loader = [(torch.tensor([i]), {"label": b["label"]}) for i, b in enumerate(batches)]

    def wrapper(x):                              # the reference methods clamp their inputs in place -> hand out copies
        b = batches[int(x[0])]
        return {k: v.clone() for k, v in b.items() if k != "label"}

    dev = torch.device("cpu")
    with torch.no_grad():
        out["full_vertices0"] = sn.feat_to_full_vertices(batches[0]["ingredients"], batches[0]["attn_cls"].clone()).numpy().copy()
        cv = init_mod.init_class_vertices(loader, wrapper, graph=sn, device=dev)
        out["class_vertices_full"] = cv.numpy().copy()
        init_w, valid = cv.topk(Vc, dim=1)
        sn.register_class_vertices(valid)
        sn.vertex_weights.copy_(init_w)
        out["valid_vertices"], out["vertex_weights"] = valid.numpy().copy(), init_w.numpy().copy()
        out["edge_weights_init"] = sn.edge_weights.tensor.detach().numpy().copy()
        out["limited_edges0"] = sn.feat_to_limited_edges(batches[0]["ingredients"], batches[0]["attn"].clone(),
                                                         batches[0]["label"]).numpy().copy()
        init_mod.init_graph(loader, wrapper, graph=sn, device=dev)
        out["edge_weights_final"] = sn.edge_weights.tensor.detach().numpy().copy()
        out["vertex_weights_final"] = sn.vertex_weights.tensor.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: nb={nb} B={B} M={M} K={K} Vc={Vc}")


def run_train_case(ns, name, B, d, M, K, Vc, D, seed):
    """One training iteration of the head as SchemaNetTrainer.train_iter runs it (tasks/worker_schema_net.py:120-140):
    normalize(), forward (SchemaNet.forward, get_atlas, Matcher.forward in grad mode), SchemaInferenceLoss
    (loss/schema_inference_loss.py:21-47) with the shipped weights (config/*/schema_net/*.yaml: cls 1, re_entropy_vertex 0.5,
    re_entropy_edge 0.75), backward.  Stores the loss terms and the gradient of every parameter."""
    from schema_inference.loss.schema_inference_loss import SchemaInferenceLoss
    vocab, mid, attn, attn_cls = ho.synth_inputs(B, d, M, seed, L, "easy")
    schema = ho.synth_schema(M, K, Vc, seed + 1)
    schema["w_v"] = torch.tensor([[0.25], [1.5]])
    schema["w_e"] = torch.tensor([[2.0], [0.125]])
    schema["vertex_weights"][:, 2::5] = 0.0                    # pruned vertices (normalised weight 1e-5 / sum <= 0.001)
    schema["vertex_weights"][1, :3] = 0.0
    gnn = ho.synth_gnn(M, D, seed + 2)
    g = torch.Generator().manual_seed(seed + 3)
    gnn["layers.0.norm.weight"] = torch.rand(D, generator=g) + 0.5
    gnn["layers.1.norm.bias"] = 0.1 * torch.randn(D, generator=g)
    gnn["fc.bias"] = 0.1 * torch.randn(D, generator=g)
    label = torch.randint(0, K, (B,), generator=g)
    disc, sn, matcher = load_params(ns, M, d, K, Vc, D, vocab, schema, gnn)
    out = dict(vocab=vocab.numpy(), mid_feat=mid.numpy(), attn=attn.numpy(), attn_cls=attn_cls.numpy(), label=label.numpy(),
               cfg=np.array([B, d, M, K, Vc, D], dtype=np.int64))
    for k, v in schema.items():
        out["schema." + k] = v.numpy()
    for k, v in gnn.items():
        out["gnn." + k] = v.numpy()
    with torch.no_grad():
        ad = ns.Adapter()
        _, match = disc(ad.adapt(mid))
        ingredients = match.t().contiguous()
    out["ingredients"] = ingredients.numpy()
    sn.train(); matcher.train()
    sn.normalize()
    for k, v in sn.state_dict().items():
        out["after_normalize." + k] = v.numpy().copy()
    inst = sn(ingredients, attn.clone(), attn_cls.clone())
    atlas = sn.get_atlas()
    out["pruned_entries"] = np.array([(torch.from_numpy(out["after_normalize.edge_weights.tensor"]) != sn.edge_weights.tensor).sum().item()],
                                     dtype=np.int64)
    pred = matcher(inst, atlas)
    loss_fn = SchemaInferenceLoss(re_a_vertex=3.0, re_a_edge=4.0)
    terms = loss_fn({"pred": pred, **atlas}, {"label": label})
    weights = {"cls": 1.0, "re_entropy_vertex": 0.5, "re_entropy_edge": 0.75}
    total = sum(terms[k] * w for k, w in weights.items())
    total.backward()
    out["pred"] = pred.detach().numpy()
    out["loss_total"] = np.array([total.item()], dtype=np.float64)
    for k, v in terms.items():
        out["loss." + k] = np.array([v.item()], dtype=np.float64)
    for prefix, mod in (("grad.schema_net.", sn), ("grad.matcher.", matcher)):
        for k, p_ in mod.named_parameters():
            if p_.requires_grad:
                out[prefix + k] = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).numpy().copy()
    out["after_step.edge_weights.tensor"] = sn.edge_weights.tensor.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: pruned {out['pruned_entries'][0]} loss {total.item():.6f} terms { {k: round(v.item(), 5) for k, v in terms.items()} }")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)          # fixtures must not depend on the thread count of the generating host
    ns = ref_harness.import_reference()
    run_head_case(ns, "head_tiny_easy", B=4, d=64, M=128, K=6, Vc=128, D=32, seed=101, mode="easy")
    run_head_case(ns, "head_hard_edge", B=4, d=48, M=256, K=4, Vc=100, D=64, seed=202, mode="hard",
                  edit=edit_edge_cases, w_v=[0.25, 1.5], w_e=[2.0, 0.125])
    run_head_case(ns, "head_wide", B=2, d=192, M=1024, K=3, Vc=64, D=256, seed=303, mode="easy")
    run_init_case(ns, "init_apis", B=3, M=64, K=4, Vc=24, seed=404)
    run_init_callers_case(ns, "init_callers", nb=2, B=5, M=48, K=3, Vc=16, seed=505)
    run_train_case(ns, "train_step", B=3, d=32, M=40, K=3, Vc=24, D=32, seed=606)


if __name__ == "__main__":
    main()
