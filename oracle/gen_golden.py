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


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)          # fixtures must not depend on the thread count of the generating host
    ns = ref_harness.import_reference()
    run_head_case(ns, "head_tiny_easy", B=4, d=64, M=128, K=6, Vc=128, D=32, seed=101, mode="easy")
    run_head_case(ns, "head_hard_edge", B=4, d=48, M=256, K=4, Vc=100, D=64, seed=202, mode="hard",
                  edit=edit_edge_cases, w_v=[0.25, 1.5], w_e=[2.0, 0.125])
    run_head_case(ns, "head_wide", B=2, d=192, M=1024, K=3, Vc=64, D=256, seed=303, mode="easy")
    run_init_case(ns, "init_apis", B=3, M=64, K=4, Vc=24, seed=404)


if __name__ == "__main__":
    main()
