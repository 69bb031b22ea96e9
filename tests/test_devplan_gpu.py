"""Device-side plan construction (stargcn_b200.devgraph) against the host mirror of
``StackedHeterGCNLayers.gen_plan`` (mxgraph/layers/layers.py:260-337): node lists, local indices, restore indices
and every per-level CSR are BIT-EXACT, on the small two-type graph of tests/test_plan_cpu.py (ids that are not
0..N-1, duplicate requests) and on the ML-100k shape; and the whole model gives the same loss and gradients whether
its plans are built on the host or on the device."""
import numpy as np
import pytest
import torch

from hostgraph import HostCSR, HostGraph, from_synth

pytestmark = pytest.mark.gpu


def host(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def small_graphs(seed=0, n_user=12, n_item=9, nnz=40, R=3):
    from stargcn_b200.devgraph import DeviceCSRMat, DeviceHeterGraph
    from stargcn_b200.sampler import DeviceCSR
    rs = np.random.RandomState(seed)
    flat = np.sort(rs.choice(n_user * n_item, nnz, replace=False))
    u, i = flat // n_item, flat % n_item
    levels = np.arange(1, R + 1).astype(np.float32)
    vals = levels[rs.randint(0, R, nnz)]
    sup = rs.uniform(0.1, 1.0, nnz).astype(np.float32)
    uid, iid = np.arange(100, 100 + n_user, dtype=np.int32), np.arange(500, 500 + n_item, dtype=np.int32)

    def csr(r, c, n_r):
        order = np.lexsort((c, r))
        ptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=n_r))]).astype(np.int32)
        return ptr, c[order].astype(np.int32), vals[order], sup[order]
    pu, cu, vu, su = csr(u, i, n_user)
    pi, ci, vi, si = csr(i, u, n_item)
    hg = HostGraph({("user", "item"): HostCSR(pu, cu, vu, levels, uid, iid, su),
                    ("item", "user"): HostCSR(pi, ci, vi, levels, iid, uid, si)})
    dg = DeviceHeterGraph(hg.meta_graph, {
        ("user", "item"): DeviceCSRMat(DeviceCSR(pu, cu, vu, levels, n_item, support=su), uid, iid),
        ("item", "user"): DeviceCSRMat(DeviceCSR(pi, ci, vi, levels, n_user, support=si), iid, uid)})
    return hg, dg, R


def stack(meta_graph, R, depth):
    from stargcn_b200.layers import HeterGCNLayer, StackedHeterGCNLayers
    mls = {("user", "item"): R, ("item", "user"): R}
    enc = StackedHeterGCNLayers()
    for _ in range(depth):
        enc.add(HeterGCNLayer(meta_graph=meta_graph, multi_link_structure=mls, agg_units=12, out_units=8,
                              agg_accum="sum", agg_act="leaky", out_act="leaky"))
    return enc


def assert_plans_equal(host_plan, dev_plan):
    (req_h, plan_h), (req_d, plan_d) = host_plan, dev_plan
    assert set(req_h) == set(req_d)
    for key in req_h:
        assert np.array_equal(req_h[key], host(req_d[key]))
    assert len(plan_h) == len(plan_d)
    for (ids_h, args_h), (ids_d, args_d) in zip(plan_h, plan_d):
        assert set(ids_h) == set(ids_d) and set(args_h) == set(args_d)
        for key in ids_h:
            assert np.array_equal(ids_h[key], host(ids_d[key])), key
        for src in args_h:
            rows_h, restore_h, ent_h = args_h[src]
            rows_d, restore_d, ent_d = args_d[src]
            assert np.array_equal(rows_h, host(rows_d))
            assert (restore_h is None) == (restore_d is None)
            if restore_h is not None:
                assert np.array_equal(restore_h, host(restore_d))
            assert set(ent_h) == set(ent_d)
            for dst in ent_h:
                ep_l, _vals, ptr_l, sup_l = ent_h[dst][:4]
                ep_d, ptr_d, sup_d = ent_d[dst][0].to_lists()
                assert ent_d[dst][0].n_nb == len(ids_h[dst])
                for r in range(len(ptr_l)):
                    n = int(ptr_l[r][-1])
                    assert np.array_equal(ptr_l[r], ptr_d[r]), (src, dst, r)
                    assert np.array_equal(np.asarray(ep_l[r][:n], np.int32), ep_d[r]), (src, dst, r)
                    assert np.array_equal(np.asarray(sup_l[r][:n], np.float32), sup_d[r]), (src, dst, r)


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_device_gen_plan_bit_exact_small(depth):
    from stargcn_b200 import devgraph
    hg, dg, R = small_graphs()
    enc = stack(hg.meta_graph, R, depth)
    sel = {"user": np.array([103, 101, 103, 110], np.int32), "item": np.array([505, 505, 500], np.int32)}
    fan = {("user", "item"): -1, ("item", "user"): -1}
    assert_plans_equal(enc.gen_plan(hg, sel, graph_sampler_args=fan, symm=True),
                       devgraph.gen_plan(enc, dg, sel, graph_sampler_args=fan, symm=True))
    # one node type only requested: the other type still appears through the neighbourhoods
    sel1 = {"user": np.array([111, 100], np.int32)}
    assert_plans_equal(enc.gen_plan(hg, sel1, graph_sampler_args=fan, symm=True),
                       devgraph.gen_plan(enc, dg, sel1, graph_sampler_args=fan, symm=True))


def test_device_gen_plan_bit_exact_ml100k_and_merge():
    from stargcn_b200 import devgraph, synth
    from stargcn_b200.hetergraph import merge_node_ids_dict
    g = synth.make_bipartite(*synth.SHAPES["ml-100k"][:3], 5, seed=1000)
    hg, dg = from_synth(g), devgraph.DeviceHeterGraph.from_synth(g)
    enc = stack(hg.meta_graph, 5, 2)
    rs = np.random.RandomState(1)
    pick = rs.choice(g["nnz"], 3000, replace=False)
    reqs = [{"user": g["u2i"]["rows"][pick].astype(np.int32), "item": g["u2i"]["cols"][pick].astype(np.int32)},
            {"user": rs.permutation(g["n_user"])[:90].astype(np.int32), "item": rs.permutation(g["n_item"])[:160].astype(np.int32)},
            {}]
    sel_h, idx_h = merge_node_ids_dict(reqs)
    sel_d, idx_d = devgraph.merge_node_ids_dict(reqs, "cuda")
    for key in sel_h:
        assert np.array_equal(sel_h[key], host(sel_d[key]))
    for a, b in zip(idx_h, idx_d):
        assert set(a) == set(b) and all(np.array_equal(a[k], host(b[k])) for k in a)
    fan = {("user", "item"): -1, ("item", "user"): -1}
    assert_plans_equal(enc.gen_plan(hg, sel_h, graph_sampler_args=fan, symm=True),
                       devgraph.gen_plan(enc, dg, sel_d, graph_sampler_args=fan, symm=True))


def test_model_same_result_with_host_and_device_plans():
    from stargcn_b200 import devgraph, synth
    from stargcn_b200.model import StarGCN
    R, D = 5, 64
    g = synth.make_bipartite(*synth.SHAPES["ml-100k"][:3], R, seed=1000)
    hg, dg = from_synth(g), devgraph.DeviceHeterGraph.from_synth(g)
    rs = np.random.RandomState(0)
    pick = rs.choice(g["nnz"], 2000, replace=False)
    pairs = np.stack([g["u2i"]["rows"][pick], g["u2i"]["cols"][pick]]).astype(np.int32)
    ratings = torch.from_numpy(g["u2i"]["vals"][pick].astype(np.float32)).cuda()
    noise = {"user": np.arange(g["n_user"], dtype=np.int32), "item": np.arange(g["n_item"], dtype=np.int32)}
    noise["user"][rs.permutation(g["n_user"])[:90]] = -1
    recon = {"user": rs.permutation(g["n_user"])[:90].astype(np.int32), "item": rs.permutation(g["n_item"])[:160].astype(np.int32)}
    torch.manual_seed(0)
    mls = {("user", "item"): R, ("item", "user"): R}
    model = StarGCN(hg.meta_graph, mls, {"user": g["n_user"], "item": g["n_item"]}, "user", "item", embed_units=D,
                    agg_units=250, out_units=75, n_blocks=2, mid_map=64, agg_accum="sum", act="leaky").cuda()
    fan = {("user", "item"): -1, ("item", "user"): -1}
    results = []
    for graph in (hg, dg):
        model.zero_grad(set_to_none=True)
        pr, pe, gt = model(graph, pairs, noise, recon, fan)
        loss = model.loss(pr, pe, gt, ratings, 3.5, 1.1, 0.1)
        loss.backward()
        results.append((loss.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    (l_h, g_h), (l_d, g_d) = results
    assert torch.equal(l_h, l_d)                       # identical plans -> identical kernels -> identical bits
    assert set(g_h) == set(g_d) and all(torch.equal(g_h[n], g_d[n]) for n in g_h)
