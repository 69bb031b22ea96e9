"""StaticGraphStep (one CUDA graph per training iteration on static whole-graph plans, batch edges masked instead
of removed) against StarGCN.forward on the graph with the batch edges REALLY removed (DeviceCSR.remove_edges, which
is bit-exact against the reference's remove_edges + get_support, tests/test_sampler_gpu.py), plans built by the
device gen_plan: same loss and same parameter gradients to fp32 rounding; the masked weights are bit-exact."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
R, D = 5, 64


def setup(shape="ml-100k", B=2000, act="leaky", seed=0):
    from stargcn_b200 import devgraph, synth
    from stargcn_b200.model import StarGCN
    n_user, n_item, n_edges, _, _ = synth.SHAPES[shape]
    g = synth.make_bipartite(n_user, n_item, n_edges, R, seed=1000)
    dg = devgraph.DeviceHeterGraph.from_synth(g)
    rs = np.random.RandomState(seed)
    pick = rs.choice(g["nnz"], B, replace=False)
    pairs = np.stack([g["u2i"]["rows"][pick], g["u2i"]["cols"][pick]]).astype(np.int32)
    ratings = g["u2i"]["vals"][pick].astype(np.float32)
    noise = {"user": np.arange(n_user, dtype=np.int32), "item": np.arange(n_item, dtype=np.int32)}
    noise["user"][rs.permutation(n_user)[:n_user // 20]] = -1
    noise["item"][rs.permutation(n_item)[:n_item // 20]] = -1
    recon = {"user": rs.permutation(n_user)[:n_user // 10].astype(np.int32), "item": rs.permutation(n_item)[:n_item // 10].astype(np.int32)}
    torch.manual_seed(0)
    mls = {("user", "item"): R, ("item", "user"): R}
    model = StarGCN(dg.meta_graph, mls, {"user": n_user, "item": n_item}, "user", "item", embed_units=D, agg_units=250,
                    out_units=75, n_blocks=2, mid_map=64, agg_accum="sum", act=act).cuda()
    return g, dg, model, pairs, ratings, noise, recon


def removed_graph(dg, pairs):
    """The graph the reference builds for the iteration: batch edges removed from both directions, supports from the
    new degrees (remove_edges_by_id, graph.py:952-974)."""
    from stargcn_b200.devgraph import DeviceCSRMat, DeviceHeterGraph
    ui, iu = dg["user", "item"], dg["item", "user"]
    ru, ci = ui.rows_of(torch.from_numpy(pairs[0]).cuda()), ui.cols_of(torch.from_numpy(pairs[1]).cuda())
    new_ui = ui.csr.remove_edges(ru, ci)
    new_iu = iu.csr.remove_edges(ci, ru)
    return DeviceHeterGraph(dg.meta_graph, {("user", "item"): DeviceCSRMat(new_ui, ui.row_ids, ui.col_ids),
                                            ("item", "user"): DeviceCSRMat(new_iu, iu.row_ids, iu.col_ids)})


@pytest.mark.parametrize("shape,act", [("ml-100k", "leaky"), ("douban", "identity")])
def test_static_step_matches_model_on_edge_removed_graph(shape, act):
    from stargcn_b200.static_step import StaticGraphStep
    g, dg, model, pairs, ratings, noise, recon = setup(shape, act=act)
    mean, std, lam = float(ratings.mean()), float(ratings.std()), 0.1
    fan = {("user", "item"): -1, ("item", "user"): -1}
    def reference():
        """Loss (host float) and gradient copies; nothing of its autograd graph survives the call — a retained
        graph would keep AccumulateGrad nodes bound to this stream and break the capture below."""
        reduced = removed_graph(dg, pairs)
        model(reduced, pairs, noise, recon, fan)                  # materialise the deferred shapes
        model.zero_grad(set_to_none=True)
        pr, pe, gt = model(reduced, pairs, noise, recon, fan)
        loss = model.loss(pr, pe, gt, torch.from_numpy(ratings).cuda(), mean, std, lam)
        loss.backward()
        return float(loss.detach()), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}, reduced

    ref_loss, ref_g, reduced = reference()

    step = StaticGraphStep(model, dg, pairs.shape[1], {k: len(v) for k, v in recon.items()}, rating_mean=mean, rating_std=std,
                           recon_lambda=lam)
    loss_eager = step(pairs, ratings, noise, recon, eager=True).clone()
    # the masked weights equal the weights of the really reduced graph at the kept edges and are 0 at the removed ones
    d = step.dirs[("user", "item")]
    keep = d.ws.view(torch.int32)[:d.g.nnz][d.base_pos.long()] != 0
    assert int((~keep).sum()) == pairs.shape[1]
    sup_kept = d.csr.support[keep]
    red = reduced["user", "item"].csr
    red_plan = red.sample_neighbors(None, -1)
    assert torch.equal(sup_kept, red_plan.support) and float(d.csr.support[~keep].abs().max()) == 0.0
    assert abs(float(loss_eager) - ref_loss) <= 1e-5 * abs(ref_loss)
    bad = []
    for n, p in model.named_parameters():
        if n in ref_g:
            e = rel_err(p.grad.cpu().numpy(), ref_g[n].cpu().numpy())
            if e > 2e-5:
                bad.append((n, e))
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
    assert not bad, bad

    # captured: same numbers as the eager body, and it follows new inputs
    step.capture()
    loss_graph = step(pairs, ratings, noise, recon).clone()
    torch.cuda.synchronize()
    assert torch.equal(loss_graph, loss_eager)
    rs = np.random.RandomState(9)
    pick2 = rs.choice(g["nnz"], pairs.shape[1], replace=False)
    pairs2 = np.stack([g["u2i"]["rows"][pick2], g["u2i"]["cols"][pick2]]).astype(np.int32)
    ratings2 = g["u2i"]["vals"][pick2].astype(np.float32)
    loss2 = step(pairs2, ratings2, noise, recon).clone()
    torch.cuda.synchronize()
    reduced2 = removed_graph(dg, pairs2)
    with torch.no_grad():
        pr, pe, gt = model(reduced2, pairs2, noise, recon, fan)
        want2 = model.loss(pr, pe, gt, torch.from_numpy(ratings2).cuda(), mean, std, lam)
    assert abs(float(loss2) - float(want2)) <= 1e-5 * abs(float(want2))
    assert abs(float(loss2) - float(loss_graph)) > 1e-6 * abs(float(loss_graph))     # it really is a different batch
