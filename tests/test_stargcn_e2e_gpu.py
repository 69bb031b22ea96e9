"""BASELINE.json configs[1] and [4]: ML-100k-shaped transductive and Douban-shaped inductive (user cold-start) full STAR-GCN — two stacked encoder blocks, masked
input embeddings, reconstruction decoder and rating head, D=64 — assembled by stargcn_b200.model.StarGCN,
against a torch-CPU fp64 execution of the SAME plans with plain dense/index ops in the reference's operator
order (per level FullyConnected, weighted segment sum, add_n, LeakyReLU; Dense; take; losses).

The loss agrees to 1e-5.  Every parameter gradient is held to 2e-5 + 2x the fp32 reference execution's own
distance from the fp64 answer; LeakyReLU branches inside the fp32 rounding band around 0 are taken from the
device forward (the same pinning as tests/test_layers_gpu.py:pin_kink), nothing else is relaxed."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from hostgraph import from_synth

pytestmark = pytest.mark.gpu
R, D, U, O, DM = 5, 64, 250, 75, 64


def vec_err(got, want):
    """(relative L2 error, cosine) of two flat fp64 vectors."""
    l2 = float((got - want).norm() / want.norm().clamp_min(1e-300))
    cos = float(torch.dot(got, want) / (got.norm() * want.norm()).clamp_min(1e-300))
    return l2, cos


class BranchLog:
    """Activation outputs of the device forward, in call order (aggregator, output Dense, first Dense of an
    embed map), recorded by forward hooks; the oracle replays them to pin LeakyReLU branches."""

    def __init__(self, model):
        self.outs, self.handles = [], []
        mods = []
        for enc in model.encoders:
            layer = enc[0]
            mods += list(layer.aggregators._mods) + list(layer._out_fcs._mods)
        for maps in model.embed_maps:
            mods += [m.l0 for m in maps._mods]
        for m in mods:
            self.handles.append(m.register_forward_hook(lambda mod, inp, out: self.outs.append(out.detach().cpu())))

    def close(self):
        for h in self.handles:
            h.remove()


def make_leaky(act, log=None, band=1e-5):
    """LeakyReLU(0.1) whose branch inside |z| <= band * max|z| follows the device forward (pin_kink)."""
    it = iter(log) if log is not None else None
    pinned = [0, 0]

    def leaky(z):
        if act != "leaky":
            if it is not None:
                next(it)
            return z
        pos = z > 0
        if it is not None:
            dev_out = next(it)
            assert dev_out.shape == z.shape, (dev_out.shape, z.shape)
            near = z.detach().abs() <= band * z.detach().abs().max()
            pos = torch.where(near, dev_out > 0, pos)
            pinned[0] += int((near & (z.detach() != 0)).sum())
            pinned[1] += z.numel()
        return torch.where(pos, z, 0.1 * z)
    leaky.pinned = pinned
    return leaky


def oracle(model, plans, lookups, needed, noise, recon_ids, gt_ratings, act, lam, dt=torch.float64, log=None):
    """CPU re-execution in ``dt``; parameters are leaf copies so that autograd yields reference gradients."""
    leaky_fn = make_leaky(act, log)

    def leaky(z, _act):
        return leaky_fn(z)
    P = {n: p.detach().to(dt).cpu().requires_grad_(True) for n, p in model.named_parameters()}
    i64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int64)

    def table(key):
        return P[f"embed_layers._mods.{list(model.embed_layers.keys()).index(key)}.weight"]

    def embed(key, ids, use_mask):
        ids = i64(ids)
        if not use_mask:
            return table(key)[ids]
        eff = i64(noise[key])[ids]
        m = (eff != -1)
        return table(key)[eff * m] * m[:, None].to(dt)

    def dense(x, prefix):
        return x @ P[prefix + ".weight"].T + P[prefix + ".bias"]

    feats = {k: embed(k, ids, True) for k, ids in needed.items()}
    pred_r, pred_e = [], []
    keys = list(model.embed_layers.keys())
    for b in range(len(plans)):
        ids_dict, args = plans[b][0]
        layer = model.encoders[b][0]
        h = {}
        for src, (row_inds, restore, entries) in args.items():
            (dst, entry), = entries.items()
            ep_l, _vals, ptr_l, sup_l = entry[:4]
            ai = list(layer.aggregators.keys()).index((src, dst))
            acc = 0
            for r in range(R):
                W, bb = P[f"encoders.{b}._blocks.0._aggregators._mods.{ai}.weight{r}"], P[f"encoders.{b}._blocks.0._aggregators._mods.{ai}.bias{r}"]
                nnz = int(ptr_l[r][-1])
                hr = feats[dst] @ W.T + bb
                seg = torch.from_numpy(np.repeat(np.arange(len(ptr_l[r]) - 1), np.diff(ptr_l[r])).astype(np.int64))
                contrib = torch.as_tensor(np.asarray(sup_l[r][:nnz]), dtype=dt)[:, None] * hr[i64(ep_l[r][:nnz])]
                acc = acc + torch.zeros(len(ptr_l[r]) - 1, U, dtype=dt).index_add(0, seg, contrib)
            oi = list(layer._out_fcs.keys()).index(src)
            out = leaky(dense(leaky(acc, act), f"encoders.{b}._blocks.0._out_fcs._mods.{oi}"), act)
            h[src] = out[i64(restore)] if restore is not None else out
        look = lookups[b]
        u = dense(h["user"][i64(look["rating"]["user"])], f"rating_user_projs.{b}")
        v = dense(h["item"][i64(look["rating"]["item"])], f"rating_item_projs.{b}")
        pred_r.append((u * v).sum(1))

        def emap(key, x):
            mi = list(model.embed_maps[b].keys()).index(key)
            pre = f"embed_maps.{b}._mods.{mi}"
            return dense(leaky(dense(x, pre + ".l0"), act), pre + ".l1")
        pred_e.append({k: emap(k, h[k][i64(idx)]) for k, idx in look["recon"].items()})
        if b < len(plans) - 1:
            feats = {k: emap(k, h[k][i64(idx)]) for k, idx in look["next"].items()}
    gt = {k: embed(k, ids, False) for k, ids in recon_ids.items()}
    y = torch.as_tensor(gt_ratings, dtype=dt)
    loss = sum((0.5 * (p - y) ** 2).mean() for p in pred_r)
    loss = loss + lam * sum(((gt[k] - pe[k]) ** 2).sum(-1).mean() for pe in pred_e for k in pe)
    loss.backward()
    if log is not None:
        assert leaky_fn.pinned[0] <= 1e-3 * max(leaky_fn.pinned[1], 1), leaky_fn.pinned   # the pinned set stays tiny
    return float(loss), {n: (p.grad.double() if p.grad is not None else None) for n, p in P.items()}


# (shape, setting): BASELINE.json configs[1] = ML-100k transductive; configs[4] = Douban-shaped inductive user
# cold-start — 20 % of the users never show their own embedding (noise = -1 -> zero vector, as the reference feeds
# held-out users, datasets.py:174-214 / iterators.py:329-351) and are the nodes to reconstruct.
SETTINGS = [("ml-100k", "transductive", "identity"), ("ml-100k", "transductive", "leaky"),
            ("douban", "inductive", "leaky")]


@pytest.mark.parametrize("shape,setting,act", SETTINGS)
def test_full_stargcn_two_blocks_with_reconstruction(shape, setting, act):
    from stargcn_b200 import synth
    from stargcn_b200.model import StarGCN
    from stargcn_b200.optim import FusedAdam
    n_user, n_item, n_edges, n_levels, _ = synth.SHAPES[shape]
    assert n_levels == R
    g = synth.make_bipartite(n_user, n_item, n_edges, R, seed=1000)
    graph = from_synth(g)
    rs = np.random.RandomState(0)
    B = 2000
    pick = rs.choice(g["nnz"], B, replace=False)
    pairs = np.stack([g["u2i"]["rows"][pick], g["u2i"]["cols"][pick]]).astype(np.int32)
    ratings = g["u2i"]["vals"][pick].astype(np.float32)
    noise, recon = {}, {}
    for key, n in (("user", n_user), ("item", n_item)):
        nz = np.arange(n, dtype=np.int32)
        perm = rs.permutation(n)
        if setting == "inductive" and key == "user":
            n_rec = n // 5
            recon[key] = perm[:n_rec].astype(np.int32)
            nz[perm[:n_rec]] = -1                                         # cold-start users: always the zero vector
        else:
            n_rec = n // 10
            recon[key] = perm[:n_rec].astype(np.int32)
            nz[perm[: n_rec // 2]] = -1                                   # masked to zero
            nz[perm[n_rec // 2: n_rec]] = rs.randint(0, n, n_rec - n_rec // 2)   # replaced by another node
        noise[key] = nz
    torch.manual_seed(0)
    mls = {("user", "item"): R, ("item", "user"): R}
    model = StarGCN(graph.meta_graph, mls, {"user": n_user, "item": n_item}, "user", "item", embed_units=D, agg_units=U,
                    out_units=O, n_blocks=2, mid_map=DM, agg_accum="sum", act=act).cuda()
    fan = {("user", "item"): -1, ("item", "user"): -1}
    mean, std, lam = float(ratings.mean()), float(ratings.std()), 0.1
    model(graph, pairs, noise, recon, fan)                            # materialises the lazily-shaped layers
    log = BranchLog(model)
    pr, pe, gt = model(graph, pairs, noise, recon, fan)
    log.close()
    assert len(pr) == 2 and len(pe) == 2 and pr[0].shape == (B, 1) and pe[1]["item"].shape == (len(recon["item"]), D)
    y = torch.from_numpy(ratings).cuda()
    loss = model.loss(pr, pe, gt, y, mean, std, lam)
    loss.backward()
    plans, lookups, needed = model.last_plans
    plans_h = [[(p[0][0], p[0][1])] for p in plans]                   # one depth per block
    target = (ratings - mean) / std
    ref_loss, ref_g = oracle(model, plans_h, lookups, needed, noise, recon, target, act, lam, log=log.outs)
    loss32, g32 = oracle(model, plans_h, lookups, needed, noise, recon, target, act, lam, dt=torch.float32, log=log.outs)
    assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss)
    # Depth compounds rounding (eight GEMM layers, each ~1e-6 from the fp64 answer on the 3xTF32 tensor-core
    # path against ~2e-7 for an fp32 FMA GEMM) and the residual pred - y cancels leading digits, so the bar for
    # the gradients of the WHOLE stack is 2e-5 plus twice the reference-order fp32 execution's own distance from
    # the fp64 answer — for every parameter, with and without LeakyReLU.
    checked, errs, bad = 0, {}, []
    for name, p in model.named_parameters():
        if p.grad is None:
            assert ref_g[name] is None or float(ref_g[name].abs().max()) == 0.0, name
            continue
        got, want = p.grad.double().cpu().reshape(-1), ref_g[name].reshape(-1)
        e_gpu = rel_err(got.numpy(), want.numpy())
        e_f32 = rel_err(g32[name].numpy().reshape(-1), want.numpy())
        errs[name] = (e_gpu, e_f32)
        checked += 1
        if not e_gpu <= 2e-5 + 2 * e_f32:
            l2, cos = vec_err(got, want)
            bad.append((name, e_gpu, e_f32, l2, cos))
    worst = sorted(errs.items(), key=lambda kv: -kv[1][0])[:5]
    print("worst gradient errors (device, fp32 oracle):", worst)
    for b in bad:
        print("OUT OF BAR:", b)
    assert not bad, bad
    assert checked >= 2 + 2 * (2 * 2 * R + 4 + 8 + 4)                 # tables + per block: agg, out_fc, maps, projs
    # one optimiser step through the multi-tensor clip + Adam with the reference's own hyper-parameters for this
    # config (LR 0.002, GRAD_CLIP 1.0: experiments/cfg/transductive_ml_100k.yml:48,54) keeps everything finite,
    # lowers the loss, and the loss at the NEW parameters again matches the fp64 re-execution.  (Adam's first step
    # moves every weight by ~lr * sign(g); at lr 1e-2 the purely linear 'identity' stack overshoots — the fp64
    # oracle shows the same 1.085 -> 1.163 — so the step size is the reference's, not an arbitrary one.)
    opt = FusedAdam(list(model.parameters()), learning_rate=2e-3)
    gnorm = opt.clip_global_norm(1.0)
    opt.step()
    assert float(gnorm) > 0 and all(torch.isfinite(p).all() for p in model.parameters())
    with torch.no_grad():
        pr2, pe2, gt2 = model(graph, pairs, noise, recon, fan)
        loss2 = model.loss(pr2, pe2, gt2, y, mean, std, lam)
    assert float(loss2) < float(loss)
    ref_loss2, _ = oracle(model, plans_h, lookups, needed, noise, recon, target, act, lam)
    assert abs(float(loss2) - ref_loss2) <= 1e-5 * abs(ref_loss2)
