"""GPU parity of the masked-embedding decoder path against oracle/layers.py
(numpy restatement of experiments/STAR-GCN.py:264-300, 226-246/441-454, 618-628, 428-438).
fp32 bar 1e-5 max-normalised, forward and every gradient; integer outputs bit-exact."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import layers as orl

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def make_noise(rs, n_all, mask_frac=0.3, swap_frac=0.1):
    """embed_noise as iterators.py:329-351 builds it: identity, -1 for masked nodes, another id for swapped."""
    noise = np.arange(n_all, dtype=np.int32)
    perm = rs.permutation(n_all)
    n_mask, n_swap = int(n_all * mask_frac), int(n_all * swap_frac)
    noise[perm[:n_mask]] = -1
    noise[perm[n_mask:n_mask + n_swap]] = rs.randint(0, n_all, n_swap)
    return noise


@pytest.mark.parametrize("D", [64, 75, 20])
@pytest.mark.parametrize("use_mask", [True, False])
def test_get_embed_forward_backward(D, use_mask):
    from stargcn_b200 import decoder
    rs = np.random.RandomState(0)
    n_all, n = 500, 1300                       # duplicates in node_ids on purpose
    table = rs.uniform(-0.1, 0.1, (n_all, D)).astype(np.float32)
    ids = rs.randint(0, n_all, n).astype(np.int32)
    noise = make_noise(rs, n_all)
    gout = rs.normal(size=(n, D)).astype(np.float32)
    t = dev(table).requires_grad_(True)
    out = decoder.get_embed(t, dev(ids), dev(noise) if use_mask else None, use_mask=use_mask)
    ref = orl.get_embed(table, ids, noise, use_mask)
    assert np.array_equal(host(out), ref)      # pure data movement: bit-exact
    out.backward(dev(gout))
    gref = orl.get_embed_backward(table.shape, ids, gout.astype(np.float64), noise, use_mask)
    assert rel_err(host(t.grad), gref) <= TOL


def test_get_embed_empty_and_errors():
    from stargcn_b200 import decoder
    t = torch.zeros(5, 8, device="cuda")
    out = decoder.get_embed(t, torch.zeros(0, dtype=torch.int32, device="cuda"), None, use_mask=False)
    assert out.shape == (0, 8)
    with pytest.raises(ValueError):
        decoder.get_embed(t, torch.zeros(3, dtype=torch.int32, device="cuda"), None, use_mask=True)
    with pytest.raises(TypeError):
        decoder.get_embed(t, torch.zeros(3, dtype=torch.int64, device="cuda"), None, use_mask=False)
    with pytest.raises(ValueError):
        decoder.get_embed(t.cpu(), torch.zeros(3, dtype=torch.int32), None, use_mask=False)


def test_take_rows_duplicates():
    from stargcn_b200 import decoder
    rs = np.random.RandomState(1)
    x = rs.normal(size=(40, 75)).astype(np.float32)
    idx = rs.randint(0, 40, 300).astype(np.int32)
    g = rs.normal(size=(300, 75)).astype(np.float32)
    xd = dev(x).requires_grad_(True)
    y = decoder.take_rows(xd, dev(idx))
    assert np.array_equal(host(y), x[idx])
    y.backward(dev(g))
    gref = np.zeros((40, 75), np.float64)
    np.add.at(gref, idx, g.astype(np.float64))
    assert rel_err(host(xd.grad), gref) <= TOL


@pytest.mark.parametrize("n,K,N,act", [(300, 75, 64, "leaky"), (1000, 250, 75, "leaky"), (257, 64, 64, None),
                                       (50, 75, 64, "relu"), (5000, 75, 64, "leaky")])
def test_fused_dense_forward_backward(n, K, N, act):
    from stargcn_b200 import decoder
    rs = np.random.RandomState(n + K)
    x = rs.normal(size=(n, K)).astype(np.float32)
    w = rs.uniform(-0.3, 0.3, (N, K)).astype(np.float32)
    b = rs.uniform(-0.3, 0.3, (N,)).astype(np.float32)
    g = rs.normal(size=(n, N)).astype(np.float32)
    xd, wd, bd = dev(x).requires_grad_(True), dev(w).requires_grad_(True), dev(b).requires_grad_(True)
    y = decoder.fused_dense(xd, wd, bd, act)
    z64 = orl.dense(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64))
    z32 = orl.dense(x, w, b)
    assert rel_err(host(y), orl.act_fwd(z64, act)) <= TOL
    assert rel_err(host(y), orl.act_fwd(z32, act)) <= TOL
    y.backward(dev(g))
    # gradient oracle with the activation branch the GPU took inside the rounding band around 0
    yh = host(y)
    near = np.abs(z64) <= 1e-5 * np.abs(z64).max()
    pre = np.where(near, np.where(yh > 0, 1e-30, -1e-30), z64)
    gz = orl.act_bwd(pre, g.astype(np.float64), act)
    assert rel_err(host(xd.grad), gz @ w.astype(np.float64)) <= TOL
    assert rel_err(host(wd.grad), gz.T @ x.astype(np.float64)) <= TOL
    assert rel_err(host(bd.grad), gz.sum(0)) <= TOL


def test_embed_map_recon_loss_end_to_end():
    """take -> Dense -> LeakyReLU -> Dense -> recon loss, gradients into h, the four decoder parameters and
    BOTH loss arguments (the reference does not detach gt_embeddings, STAR-GCN.py:359-363)."""
    from stargcn_b200 import decoder
    rs = np.random.RandomState(5)
    n_nodes, O, D, n_rec = 400, 75, 64, 150
    h = rs.normal(size=(n_nodes, O)).astype(np.float32)
    idx = rs.choice(n_nodes, n_rec, replace=False).astype(np.int32)
    w0 = rs.uniform(-0.2, 0.2, (D, O)).astype(np.float32); b0 = rs.uniform(-0.1, 0.1, D).astype(np.float32)
    w1 = rs.uniform(-0.2, 0.2, (D, D)).astype(np.float32); b1 = rs.uniform(-0.1, 0.1, D).astype(np.float32)
    table = rs.uniform(-0.1, 0.1, (n_nodes, D)).astype(np.float32)
    rec_ids = rs.choice(n_nodes, n_rec, replace=False).astype(np.int32)

    em = decoder.EmbedMap(D, act="leaky", in_units=O).cuda()
    with torch.no_grad():
        em.l0.weight.copy_(dev(w0)); em.l0.bias.copy_(dev(b0)); em.l1.weight.copy_(dev(w1)); em.l1.bias.copy_(dev(b1))
    hd = dev(h).requires_grad_(True)
    td = dev(table).requires_grad_(True)
    gt = decoder.get_embed(td, dev(rec_ids), None, use_mask=False)
    pred = em(hd, dev(idx))
    loss = decoder.recon_loss(gt, pred)
    loss.backward()

    # oracle in fp64 (exact) with torch autograd on the CPU restatement
    t64 = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    H, W0, B0, W1, B1, T = t64(h), t64(w0), t64(b0), t64(w1), t64(b1), t64(table)
    z = H[idx.astype(np.int64)] @ W0.T + B0
    p = torch.where(z > 0, z, 0.1 * z) @ W1.T + B1
    l_ref = ((T[rec_ids.astype(np.int64)] - p) ** 2).sum(-1).mean()
    l_ref.backward()
    assert rel_err(host(pred), orl.embed_map(h.astype(np.float64), idx, w0, b0, w1, b1)) <= TOL
    assert abs(float(loss) - float(orl.recon_loss(host(gt).astype(np.float64), p.detach().numpy()))) <= TOL * abs(float(l_ref))
    assert abs(float(loss) - float(l_ref)) <= TOL * abs(float(l_ref))
    for got, ref in ((hd.grad, H.grad), (em.l0.weight.grad, W0.grad), (em.l0.bias.grad, B0.grad),
                     (em.l1.weight.grad, W1.grad), (em.l1.bias.grad, B1.grad), (td.grad, T.grad)):
        assert rel_err(host(got), ref.numpy()) <= TOL


def test_rating_head_and_l2_loss():
    from stargcn_b200 import decoder
    from stargcn_b200.layers import InnerProductLayer
    from stargcn_b200.layers.common import Dense
    rs = np.random.RandomState(6)
    nu, ni, O, Dm, B = 120, 90, 75, 64, 1000
    hu = rs.normal(size=(nu, O)).astype(np.float32); hi = rs.normal(size=(ni, O)).astype(np.float32)
    iu = rs.randint(0, nu, B).astype(np.int32); ii = rs.randint(0, ni, B).astype(np.int32)
    wu = rs.uniform(-0.2, 0.2, (Dm, O)).astype(np.float32); bu = rs.uniform(-0.1, 0.1, Dm).astype(np.float32)
    wi = rs.uniform(-0.2, 0.2, (Dm, O)).astype(np.float32); bi = rs.uniform(-0.1, 0.1, Dm).astype(np.float32)
    label = rs.normal(size=B).astype(np.float32)
    pu, pi = Dense(Dm, in_units=O).cuda(), Dense(Dm, in_units=O).cuda()
    with torch.no_grad():
        pu.weight.copy_(dev(wu)); pu.bias.copy_(dev(bu)); pi.weight.copy_(dev(wi)); pi.bias.copy_(dev(bi))
    hud, hid = dev(hu).requires_grad_(True), dev(hi).requires_grad_(True)
    pred = InnerProductLayer()(pu(decoder.take_rows(hud, dev(iu))), pi(decoder.take_rows(hid, dev(ii))))
    assert pred.shape == (B, 1)
    loss = decoder.l2_loss(pred, dev(label))
    loss.backward()
    f64 = lambda a: a.astype(np.float64)
    ref_pred = orl.rating_head(f64(hu), f64(hi), iu, ii, f64(wu), f64(bu), f64(wi), f64(bi))
    assert rel_err(host(pred), ref_pred) <= TOL
    assert abs(float(loss) - float(orl.l2_loss(ref_pred, f64(label)))) <= TOL * float(orl.l2_loss(ref_pred, f64(label)))
    t64 = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    HU, HI, WU, WI = t64(hu), t64(hi), t64(wu), t64(wi)
    u = HU[iu.astype(np.int64)] @ WU.T + torch.tensor(f64(bu)); v = HI[ii.astype(np.int64)] @ WI.T + torch.tensor(f64(bi))
    l = (0.5 * ((u * v).sum(1) - torch.tensor(f64(label))) ** 2).mean()
    l.backward()
    assert rel_err(host(hud.grad), HU.grad.numpy()) <= TOL
    assert rel_err(host(hid.grad), HI.grad.numpy()) <= TOL
    assert rel_err(host(pu.weight.grad), WU.grad.numpy()) <= TOL
    assert rel_err(host(pi.weight.grad), WI.grad.numpy()) <= TOL


def test_losses_are_bit_reproducible():
    from stargcn_b200 import decoder
    a = torch.randn(7000, 64, device="cuda"); b = torch.randn(7000, 64, device="cuda")
    l1, l2 = decoder.recon_loss(a, b), decoder.recon_loss(a, b)
    assert torch.equal(l1, l2)
    ref = ((a.double() - b.double()) ** 2).sum(-1).mean()
    assert abs(float(l1) - float(ref)) <= 1e-6 * float(ref)
