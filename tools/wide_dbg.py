import sys, os
sys.path.insert(0, 'schemanet-pytorch_b200'); sys.path.insert(0, 'oracle')
import torch, head_oracle as ho
from schemanet_b200 import native
from schema_inference.graph import GNN, Matcher

def rel(a, b):
    return float((a.cpu().double() - b.double()).abs().max() / b.double().abs().max())

def class_case(K, Vc, D, M, thr):
    sch = ho.synth_schema(M, K, Vc, seed=32)
    params = ho.synth_gnn(M, D, seed=31)
    gnn = GNN(M, D, num_layers=2).cuda(); gnn.load_state_dict(params)
    _, _, f = native.class_side(gnn.param_pack(), sch["vertex_weights"].cuda(), sch["edge_weights"].clone().cuda(),
                                sch["class_ingredients"].cuda(), thr, True, False, want_edges=False)
    atlas = ho.class_atlas(sch["vertex_weights"], sch["edge_weights"].clone(), sch["class_ingredients"], thr, False)
    want = ho.gnn_forward(params, atlas["class_vertices"], atlas["class_edges"], sch["class_ingredients"], None)
    nact = (atlas["class_vertices"] > (thr or -1)).sum(1).tolist()
    print(f"class K={K} Vc={Vc} D={D} M={M} thr={thr}: rel {rel(f, want):.2e}  n_act {nact[:4]}")

def generic_case(G, n, D, M):
    gen = torch.Generator().manual_seed(1)
    params = ho.synth_gnn(M, D, seed=31)
    m = Matcher("inner_product", M, dict(embed_dim=D, num_layers=2)).cuda(); m.gnn.load_state_dict(params)
    nodes = torch.rand(G, n, generator=gen) / n; edges = torch.rand(G, n, n, generator=gen) / n
    ids = torch.randint(0, M, (G, n), generator=gen)
    want = ho.gnn_forward(params, nodes, edges, ids, None)
    with torch.no_grad():
        got = m.gnn(nodes.cuda(), edges.cuda(), ids.cuda())
    print(f"generic G={G} n={n} D={D} M={M}: rel {rel(got, want):.2e}")

generic_case(40, 300, 1024, 700)
generic_case(24, 500, 1024, 1200)
class_case(3, 300, 512, 320, None)
class_case(3, 300, 512, 320, 0.001)
class_case(3, 300, 512, 320, 0.0034)
class_case(24, 500, 1024, 1200, 0.001)
class_case(24, 500, 1024, 1200, None)
os.environ["SCHEMANET_WIDE_UNFUSED"] = "1"
