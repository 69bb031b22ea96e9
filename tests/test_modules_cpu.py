"""Host-side behaviour of the layer mirror that needs no device: deferred-shape parameters follow torch's
lazy-module protocol (state_dict / load_state_dict on a freshly built model — the counterpart of
gluon's load_parameters on a deferred-init block), constructor surface of the reference classes."""
import pytest
import torch

import stargcn_b200  # noqa: F401
from stargcn_b200.layers import GCNAggregator, HeterGCNLayer, MultiLinkGCNAggregator, StackedHeterGCNLayers
from stargcn_b200.layers.common import Dense
from stargcn_b200.model import StarGCN

META = {"user": {"item": "rating"}, "item": {"user": "rev_rating"}}
MLS = {("user", "item"): 5, ("item", "user"): 5}


def test_dense_state_dict_before_first_forward():
    d = Dense(8)
    sd = d.state_dict()                                   # used to raise "uninitialized parameter"
    assert set(sd) == {"weight", "bias"}                  # bias is registered up front
    src = Dense(8, in_units=5)
    d.load_state_dict(src.state_dict())                   # materialises from the checkpoint's shapes
    assert d.weight.shape == (8, 5) and torch.equal(d.weight, src.weight) and torch.equal(d.bias, src.bias)
    assert not d.has_uninitialized_params()
    nob = Dense(4, use_bias=False)
    assert set(nob.state_dict()) == {"weight"}


def test_aggregator_state_dict_round_trip():
    a = MultiLinkGCNAggregator(units=6, num_links=3, accum="sum")
    assert len(a.state_dict()) == 6
    src = MultiLinkGCNAggregator(units=6, num_links=3, accum="sum", in_units=4)
    a.load_state_dict(src.state_dict())
    assert a.weight2.shape == (6, 4) and torch.equal(a.bias1, src.bias1)
    g = GCNAggregator(units=6)
    assert sorted(g.state_dict()) == ["_agg.bias0", "_agg.weight0"]


def test_whole_model_checkpoint_into_fresh_model():
    def build():
        return StarGCN(META, MLS, {"user": 30, "item": 20}, "user", "item", embed_units=16, agg_units=20, out_units=12,
                       n_blocks=2, mid_map=8, agg_accum="sum", act="leaky")
    torch.manual_seed(0)
    fresh = build()
    keys = set(fresh.state_dict())                        # works before any forward
    # a "trained" model: every deferred shape resolved the way a forward pass would resolve it
    trained = build()
    for m in trained.modules():
        if isinstance(m, MultiLinkGCNAggregator):
            m._materialize(16, None)
    for enc in trained.encoders:
        for fc in enc[0]._out_fcs._mods:
            fc._materialize(20, None)
    for maps in trained.embed_maps:
        for em in maps._mods:
            em.l0._materialize(12, None); em.l1._materialize(16, None)
    for proj in list(trained.rating_user_projs) + list(trained.rating_item_projs):
        proj._materialize(12, None)
    assert set(trained.state_dict()) == keys
    fresh.load_state_dict(trained.state_dict())
    for (n1, p1), (n2, p2) in zip(fresh.named_parameters(), trained.named_parameters()):
        assert n1 == n2 and p1.shape == p2.shape and torch.equal(p1, p2)
    assert not any(isinstance(p, torch.nn.UninitializedParameter) for p in fresh.parameters())


def test_reference_constructor_surface():
    layer = HeterGCNLayer(meta_graph=META, multi_link_structure=MLS, agg_units=20, out_units=12, source_keys=None,
                          dropout_rate=0.1, agg_ordinal_sharing=False, agg_accum="sum", agg_act="leaky", layer_accum="stack",
                          accum_self=False, out_act="leaky", prefix="gcn0_", params=None)
    assert ("user", "item") in layer.aggregators and layer.aggregators[("item", "user")].use_multi_link
    stack = StackedHeterGCNLayers(recurrent_layer_num=None, prefix="enc_")
    stack.add(layer)
    assert len(stack) == 1 and stack[0] is layer
    with pytest.raises(TypeError):
        stack.add(torch.nn.Linear(2, 2))
