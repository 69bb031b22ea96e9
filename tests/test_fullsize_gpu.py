"""BASELINE.json configs at their full sizes.

ML-1M shape (config 3): direct comparison with the reference-order oracle (finishes in seconds).
ML-10M shape (config 4, the benchmark workload): size-independent properties of the fused path —
  adjointness  <A x, g> == <x, A^T g>   (the backward gather is the exact transpose of the forward one)
  linearity    A(a x + b y) == a A x + b A y   (identity activation, zero bias)
  support sums wsum == segment sums of the support array (checked through the bias term)
  determinism  two runs are bit-identical (no float atomics anywhere)
and the CUDA-graph replay of the step equals the eager step bit for bit."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import layers as orl

pytestmark = pytest.mark.gpu
TOL = 1e-5
U = 250


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build(shape, act, seed=0, zero_bias=False):
    from stargcn_b200 import synth
    from stargcn_b200.graph import MultiLinkCSR
    from stargcn_b200.layers import MultiLinkGCNAggregator
    wl = synth.make_layer_inputs(shape, seed=1000)
    R, D = wl["R"], wl["D"]
    rs = np.random.RandomState(seed)
    bound = np.sqrt(3.0 / D)
    ws = [rs.uniform(-bound, bound, (U, D)).astype(np.float32) for _ in range(R)]
    bs = [np.zeros(U, np.float32) if zero_bias else rs.uniform(-0.1, 0.1, U).astype(np.float32) for _ in range(R)]
    ep_l, ptr_l, sup_l, _ = wl["user"]
    csr = MultiLinkCSR(ep_l, ptr_l, sup_l, n_nb=wl["n_item"], device="cuda")
    agg = MultiLinkGCNAggregator(units=U, num_links=R, act=act, ordinal_sharing=False, accum="sum", in_units=D).cuda()
    with torch.no_grad():
        for i in range(R):
            getattr(agg, f"weight{i}").copy_(dev(ws[i]))
            getattr(agg, f"bias{i}").copy_(dev(bs[i]))
    return wl, csr, agg, ws, bs


def test_ml1m_shape_vs_oracle():
    wl, csr, agg, ws, bs = build("ml-1m", "leaky")
    ep_l, ptr_l, sup_l, _ = wl["user"]
    x = wl["x_item"]
    rs = np.random.RandomState(4)
    gout = rs.normal(size=(wl["n_user"], U)).astype(np.float32)
    ref_out, pre = orl.multilink_aggregator_forward(x, ws, bs, ep_l, ptr_l, sup_l, "sum", "leaky")
    xd = dev(x).requires_grad_(True)
    out = agg(xd, csr)
    assert rel_err(out.detach().cpu().numpy(), ref_out) <= TOL
    oh = out.detach().cpu().numpy()
    near = np.abs(pre) <= 1e-5 * np.abs(pre).max()
    pre = np.where(near, np.where(oh > 0, np.abs(pre) + 1e-30, -np.abs(pre) - 1e-30), pre).astype(np.float32)
    gx_ref, gw_ref, gb_ref = orl.multilink_aggregator_backward(x, ws, bs, ep_l, ptr_l, sup_l, gout, pre, "sum", "leaky")
    out.backward(dev(gout))
    assert rel_err(xd.grad.cpu().numpy(), gx_ref) <= TOL
    assert rel_err(agg.weight1.grad.cpu().numpy(), gw_ref[1]) <= TOL
    assert rel_err(agg.bias4.grad.cpu().numpy(), gb_ref[4]) <= TOL


@pytest.fixture(scope="module")
def ml10m():
    return build("ml-10m", None, zero_bias=True)


def test_ml10m_adjoint_linearity_determinism(ml10m):
    wl, csr, agg, ws, bs = ml10m
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(wl["n_item"], wl["D"], device="cuda", generator=g).requires_grad_(True)
    y = torch.randn(wl["n_item"], wl["D"], device="cuda", generator=g)
    gout = torch.randn(wl["n_user"], U, device="cuda", generator=g)
    out = agg(x, csr)
    out.backward(gout)
    lhs = (out.detach().double() * gout.double()).sum()
    rhs = (x.detach().double() * x.grad.double()).sum()
    assert abs(float(lhs - rhs)) <= TOL * abs(float(lhs))            # <A x, g> == <x, A^T g>
    with torch.no_grad():
        comb = agg(2.0 * x.detach() - 0.5 * y, csr)
        want = 2.0 * out.detach() - 0.5 * agg(y, csr)
    assert rel_err(comb.cpu().numpy(), want.cpu().numpy()) <= TOL     # linearity
    with torch.no_grad():
        again = agg(x.detach(), csr)
    assert torch.equal(again, out.detach())                           # bit-identical reruns
    x2 = x.detach().clone().requires_grad_(True)
    agg(x2, csr).backward(gout)
    assert torch.equal(x2.grad, x.grad)


def test_ml10m_support_sums_through_bias(ml10m):
    """With x = 0 the layer output is sum_r wsum[i, r] * b_r: checks the per-segment support sums."""
    wl, csr, agg, ws, bs = ml10m
    R = wl["R"]
    with torch.no_grad():
        for i in range(R):
            getattr(agg, f"bias{i}").fill_(float(i + 1))
        out = agg(torch.zeros(wl["n_item"], wl["D"], device="cuda"), csr)
        for i in range(R):
            getattr(agg, f"bias{i}").zero_()
    ep_l, ptr_l, sup_l, _ = wl["user"]
    want = np.zeros(wl["n_user"], np.float64)
    for r in range(R):
        seg = np.add.reduceat(np.concatenate([sup_l[r].astype(np.float64), [0.0]]), np.minimum(ptr_l[r][:-1], len(sup_l[r])))
        seg[np.diff(ptr_l[r]) == 0] = 0.0
        want += (r + 1) * seg
    got = out.cpu().numpy()
    assert rel_err(got[:, 0], want) <= TOL and rel_err(got[:, U - 1], want) <= TOL


def test_graph_replay_equals_eager(ml10m):
    from stargcn_b200 import runtime
    wl, csr, agg, ws, bs = ml10m
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(wl["n_item"], wl["D"], device="cuda", generator=g).requires_grad_(True)
    gout = torch.randn(wl["n_user"], U, device="cuda", generator=g)
    holder = {}

    def step():
        x.grad = None
        for p in agg.parameters():
            p.grad = None
        out = agg(x, csr)
        out.backward(gout)
        holder["out"] = out.detach()      # no reference to the autograd graph may outlive the step (capture rule)

    step()
    eager_out, eager_gx, eager_gw = holder["out"].detach().clone(), x.grad.clone(), agg.weight3.grad.clone()
    graphed = runtime.GraphedStep(step)
    with torch.no_grad():
        x.mul_(0.5)                                   # new input values in the same buffers
    graphed()
    torch.cuda.synchronize()
    assert rel_err(holder["out"].detach().cpu().numpy(), (0.5 * eager_out).cpu().numpy()) <= TOL
    with torch.no_grad():
        x.mul_(2.0)
    graphed()
    torch.cuda.synchronize()
    assert torch.equal(holder["out"].detach(), eager_out)
    assert torch.equal(x.grad, eager_gx) and torch.equal(agg.weight3.grad, eager_gw)
