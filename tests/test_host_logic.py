"""CPU: host-side logic of the drop-in modules that needs no GPU (shapes, state-dict keys, sharding, errors)."""
import pytest
import torch

from schemanet_b200 import dist as shdist


def test_state_dict_keys_match_reference_names():
    from schema_inference.graph import SchemaNet, Matcher
    sn = SchemaNet(32, num_classes=3, class_max_vertices=16, prune_node_threshold=0.001)
    keys = set(sn.state_dict().keys())
    assert keys == {"class_ingredients.tensor", "vertex_weights.tensor", "edge_weights.tensor",
                    "vertex_attribute_weights.tensor", "edge_attribute_weights.tensor"}
    m = Matcher("inner_product", 32, dict(embed_dim=8, num_layers=2, identity_proj=False, activation="relu"))
    mk = set(m.state_dict().keys())
    want = {"gnn.embedding.weight", "gnn.fc.weight", "gnn.fc.bias"}
    for i in range(2):
        want |= {f"gnn.layers.{i}.g_conv.linear.weight", f"gnn.layers.{i}.g_conv.linear.bias",
                 f"gnn.layers.{i}.norm.weight", f"gnn.layers.{i}.norm.bias"}
    assert mk == want


def test_schema_net_init_and_registration():
    from schema_inference.graph import SchemaNet
    sn = SchemaNet(20, num_classes=2, class_max_vertices=5, constant_vertex_attr=(0.2, 0.8))
    assert not sn.vertex_attribute_weights.tensor.requires_grad and sn.edge_attribute_weights.tensor.requires_grad
    assert torch.allclose(sn.vertex_attribute_weights.tensor.flatten(), torch.tensor([0.2, 0.8]))
    assert torch.allclose(sn.vertex_weights.tensor.sum(-1), torch.ones(2), atol=1e-5)
    cv = torch.tensor([[4, 7, 1, 0, 9], [3, 3, 2, 8, 5]])
    sn.register_class_vertices(cv)
    assert sn.class_ingredient_dict[0] == {4: 0, 7: 1, 1: 2, 0: 3, 9: 4}
    assert sn.class_ingredient_dict[1][3] == 1          # later duplicate wins, like the reference's dict comprehension
    sd = sn.state_dict()
    sn2 = SchemaNet(20, num_classes=2, class_max_vertices=5)
    sn2.load_state_dict(sd)
    assert sn2.class_ingredient_dict == sn.class_ingredient_dict


def test_cpp_extension_error_behaviour():
    import cpp_extension
    ing = torch.zeros(2, 4, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="Batch size is not compat"):
        cpp_extension.cpp_feat_to_instance_e(ing, torch.zeros(2, 4, 4), torch.zeros(4, 4), [{}], torch.ones(2, 1), True)
    with pytest.raises(RuntimeError, match="diagonal dimensions"):
        cpp_extension.cpp_feat_to_instance_e(ing, torch.zeros(2, 4, 4), torch.zeros(4, 4), [{}, {}], torch.ones(2, 1),
                                             True, True)
    with pytest.raises(RuntimeError, match="Long"):
        cpp_extension.cpp_feat_to_v_attr(ing.int(), torch.zeros(2, 4), 8, True)
    assert not hasattr(cpp_extension, "__all__") or "cpp_feat_to_instance_e" not in cpp_extension.__all__


def test_shard_ranges_cover_everything():
    for n in (1, 7, 100, 101, 1000):
        for world in (1, 2, 3, 8):
            spans = [shdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_geo_table_is_cached_and_exact():
    import schema_inference.graph.utils as gu
    a = gu.pair_wise_point_sim(14, 14, 1, 2)
    assert a is gu.pair_wise_point_sim(14, 14, 1, 2)
    assert a.shape == (196, 196) and float(a[0, 0]) == 1.0 and abs(float(a[0, 1]) - 0.5) < 1e-7
