"""GPU parity: every seg_op kernel (through the C ABI) against the pinned CPU oracle and the
committed golden fixtures.  Bars: bit-exact for integer outputs (segment ids, transpose, arg-max
positions); fp32 within 1e-5 of the reference loops in the normalised error
max|a-b| / max|b| (north_star: "1e-5 relative fp32")."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import cases, segops as orc
from oracle.cases import sub

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def so():
    import stargcn_b200
    from stargcn_b200 import seg_op
    return seg_op


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("ci", range(len(cases.CONTIG_SHAPES)))
def test_contig_ops(so, golden, ci):
    b, s, n = cases.CONTIG_SHAPES[ci]
    c = cases.contig_case(100 + ci, b, s, n)
    data, rhs, ptr = dev(c["data"]), dev(c["rhs"]), dev(c["indptr"])
    assert rel_err(host(so.seg_sum(data, ptr)), orc.seg_sum(c["data"], c["indptr"])) <= TOL
    np.testing.assert_allclose(sub(host(so.seg_sum(data, ptr))), golden[f"seg_sum/{ci}/npy"], rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(host(so.seg_broadcast_add(data, rhs, ptr)), orc.seg_broadcast_add(c["data"], c["rhs"], c["indptr"]))
    np.testing.assert_array_equal(host(so.seg_broadcast_mul(data, rhs, ptr)), orc.seg_broadcast_mul(c["data"], c["rhs"], c["indptr"]))
    np.testing.assert_array_equal(host(so.seg_broadcast_to(rhs, ptr, n)), orc.seg_broadcast_to(c["rhs"], c["indptr"], n))
    sm = host(so.seg_softmax(data, ptr))
    assert rel_err(sm, orc.seg_softmax(c["data"], c["indptr"])) <= TOL
    np.testing.assert_allclose(sub(sm), golden[f"seg_softmax/{ci}/npy"], rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(host(so.seg_ids(ptr, n)), orc.seg_ids(c["indptr"], n))


def test_contig_ops_ragged_and_req(so):
    c = cases.contig_case(7, 3, 40, 500, allow_empty=True)
    data, rhs, ptr = dev(c["data"]), dev(c["rhs"]), dev(c["indptr"])
    assert rel_err(host(so.seg_sum(data, ptr)), orc.seg_sum(c["data"], c["indptr"])) <= TOL
    np.testing.assert_array_equal(host(so.seg_broadcast_to(rhs, ptr, 500)), orc.seg_broadcast_to(c["rhs"], c["indptr"], 500))
    init = np.random.RandomState(0).normal(size=(3, 40)).astype(np.float32)
    out = dev(init.copy())
    so.seg_sum(data, ptr, out=out, req="add")
    assert rel_err(host(out), orc.seg_sum(c["data"], c["indptr"], req="add", init=init)) <= TOL
    out = dev(init.copy())
    so.seg_sum(data, ptr, out=out, req="null")
    np.testing.assert_array_equal(host(out), init)
    # gradients of the contiguous ops (FGradient pairings, seg_op.cc:370-379,427-443,478-490)
    d = data.clone().requires_grad_(True)
    r = rhs.clone().requires_grad_(True)
    og = dev(c["ograd"])
    so.seg_broadcast_mul(d, r, ptr).backward(og)
    seg = orc.seg_ids(c["indptr"], 500)
    np.testing.assert_allclose(host(d.grad), c["ograd"] * c["rhs"][:, seg], rtol=1e-6, atol=1e-6)
    assert rel_err(host(r.grad), orc.seg_sum(c["ograd"] * c["data"], c["indptr"])) <= TOL
    d2 = data.clone().requires_grad_(True)
    so.seg_softmax(d2, ptr).backward(og)
    val = orc.seg_softmax(c["data"], c["indptr"])
    assert rel_err(host(d2.grad), orc.seg_softmax_bwd(c["ograd"], val, c["indptr"])) <= 1e-4


ALL_GATHER = [(200 + i, s, False) for i, s in enumerate(cases.GATHER_SHAPES)] + \
             [(300 + i, s, True) for i, s in enumerate(cases.EXTRA_GATHER_SHAPES)]


@pytest.mark.parametrize("seed,shp,ragged", ALL_GATHER)
@pytest.mark.parametrize("use_schedule", [False, True])
def test_weighted_pool_fwd_bwd(so, seed, shp, ragged, use_schedule):
    b, s, t, n, f = shp
    c = cases.gather_case(seed, *shp, allow_empty=ragged)
    pat = so.CSRPattern(dev(c["indices"]), dev(c["indptr"]), t, chunk=16, use_schedule=use_schedule)
    data = dev(c["data"]).requires_grad_(True)
    w = dev(c["weights"]).requires_grad_(True)
    out = so.seg_weighted_pool(data, w, pat.indices, pat.indptr, pattern=pat)
    ref = orc.seg_weighted_pool(c["data"], c["weights"], c["indices"], c["indptr"])
    assert out.shape == ref.shape
    assert rel_err(host(out), ref) <= TOL
    out.backward(dev(c["gout"]))
    assert rel_err(host(data.grad), orc.seg_weighted_pool_bwd_data(c["gout"], c["weights"], c["indices"], c["indptr"], t)) <= TOL
    if n:
        assert rel_err(host(w.grad), orc.seg_take_k_corr(c["gout"], c["data"], c["indices"], c["indptr"])) <= TOL
    # integer bookkeeping: bit exact
    ti, tp, ts = orc.csr_transpose(c["indices"], c["indptr"], t)
    g_ti, g_tp, g_ts = pat.transpose()
    np.testing.assert_array_equal(host(g_ti), ti)
    np.testing.assert_array_equal(host(g_tp), tp)
    np.testing.assert_array_equal(host(g_ts), ts)


@pytest.mark.parametrize("ci", range(len(cases.GATHER_SHAPES)))
def test_gather_ops_vs_golden(so, golden, ci):
    shp = cases.GATHER_SHAPES[ci]
    b, s, t, n, f = shp
    c = cases.gather_case(200 + ci, *shp)
    data, w, idx, ptr = dev(c["data"]), dev(c["weights"]), dev(c["indices"]), dev(c["indptr"])
    out = host(so.seg_weighted_pool(data, w, idx, ptr))
    assert rel_err(sub(out), golden[f"weighted_pool/{ci}/ref"]) <= TOL
    np.testing.assert_allclose(sub(out), golden[f"weighted_pool/{ci}/npy"], rtol=1e-4, atol=1e-4)
    kc = host(so.seg_take_k_corr(dev(c["embed1"]), data, idx, ptr))
    assert rel_err(sub(kc), golden[f"take_k_corr/{ci}/ref"]) <= TOL
    np.testing.assert_allclose(sub(kc), golden[f"take_k_corr/{ci}/npy"], rtol=1e-4, atol=1e-4)
    for pt in ("sum", "avg", "max"):
        d = data.clone().requires_grad_(True)
        val = so.seg_pool(d, idx, ptr, pool_type=pt)
        assert rel_err(sub(host(val)), golden[f"seg_pool_{pt}/{ci}/ref"]) <= TOL
        np.testing.assert_allclose(sub(host(val)), golden[f"seg_pool_{pt}/{ci}/npy"], rtol=1e-4, atol=1e-4)
        val.backward(dev(c["gout"]))
        assert rel_err(sub(host(d.grad)), golden[f"seg_pool_{pt}_bwd/{ci}/ref"]) <= TOL


@pytest.mark.parametrize("seed,shp,ragged", ALL_GATHER)
def test_seg_pool_all_types(so, seed, shp, ragged):
    b, s, t, n, f = shp
    c = cases.gather_case(seed, *shp, allow_empty=ragged)
    idx, ptr = dev(c["indices"]), dev(c["indptr"])
    for pt in ("sum", "avg", "max"):
        pat = so.CSRPattern(idx, ptr, t, chunk=16, use_schedule=(pt != "max"))
        d = dev(c["data"]).requires_grad_(True)
        val = so.seg_pool(d, idx, ptr, pool_type=pt, pattern=pat)
        ref_val, ref_am = orc.seg_pool(c["data"], c["indices"], c["indptr"], pt, return_argmax=True)
        if pt == "max":
            np.testing.assert_array_equal(host(val), ref_val)          # max picks an input value: exact
        else:
            assert rel_err(host(val), ref_val) <= TOL
        val.backward(dev(c["gout"]))
        ref_g = orc.seg_pool_bwd(c["gout"], ref_am if pt == "max" else None, c["indices"], c["indptr"], t, pt)
        assert rel_err(host(d.grad), ref_g) <= TOL


def test_argmax_bit_exact_and_ties(so):
    import stargcn_b200
    from stargcn_b200.seg_op import _seg_pool_fwd
    rs = np.random.RandomState(3)
    data = rs.randint(-3, 4, size=(2, 9, 6)).astype(np.float32)   # many ties
    idx = rs.randint(0, 9, size=80).astype(np.int32)
    ptr = cases.rand_indptr(rs, 12, 80, allow_empty=True)
    pat = so.CSRPattern(dev(idx), dev(ptr), 9)
    val, am = _seg_pool_fwd(dev(data), pat, "max")
    ref_val, ref_am = orc.seg_pool(data, idx, ptr, "max", return_argmax=True)
    np.testing.assert_array_equal(host(am), ref_am)
    np.testing.assert_array_equal(host(val), ref_val)


def test_req_add_and_null(so):
    shp = (2, 33, 17, 300, 64)
    c = cases.gather_case(11, *shp, allow_empty=True)
    data, w, idx, ptr = dev(c["data"]), dev(c["weights"]), dev(c["indices"]), dev(c["indptr"])
    for use_schedule in (False, True):
        pat = so.CSRPattern(idx, ptr, 17, chunk=8, use_schedule=use_schedule)
        out = dev(c["init_out"].copy())
        so.seg_weighted_pool(data, w, idx, ptr, pattern=pat, out=out, req="add")
        ref = orc.seg_weighted_pool(c["data"], c["weights"], c["indices"], c["indptr"], req="add", init=c["init_out"])
        assert rel_err(host(out), ref) <= TOL
        out = dev(c["init_out"].copy())
        so.seg_weighted_pool(data, w, idx, ptr, pattern=pat, out=out, req="null")
        np.testing.assert_array_equal(host(out), c["init_out"])


def test_errors(so):
    c = cases.gather_case(1, 1, 5, 10, 30, 8)
    data, w, idx, ptr = dev(c["data"]), dev(c["weights"]), dev(c["indices"]), dev(c["indptr"])
    with pytest.raises(TypeError):
        so.seg_weighted_pool(data.double(), w, idx, ptr)
    with pytest.raises(TypeError):
        so.seg_weighted_pool(data, w, idx.long(), ptr)
    with pytest.raises(ValueError):
        so.seg_weighted_pool(data.cpu(), w, idx, ptr)
    with pytest.raises(ValueError):
        so.seg_weighted_pool(data, w[:, :-1], idx, ptr)
    with pytest.raises(ValueError):
        so.seg_pool(data, idx, ptr, pool_type="median")


def test_large_heavy_tailed_properties(so):
    """BASELINE-size property checks the oracle is too slow for: linearity in the weights and
    equality of the scheduled (split) and unscheduled kernels on a heavy-tailed pattern."""
    rs = np.random.RandomState(5)
    n_seg, n_nb, F = 20000, 5000, 64
    lens = np.minimum((rs.pareto(1.1, n_seg) * 20).astype(np.int64), 30000)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nnz = int(ptr[-1])
    idx = rs.randint(0, n_nb, nnz).astype(np.int32)
    data = dev(rs.normal(size=(1, n_nb, F)).astype(np.float32))
    w1, w2 = dev(rs.normal(size=(1, nnz)).astype(np.float32)), dev(rs.normal(size=(1, nnz)).astype(np.float32))
    p_s = so.CSRPattern(dev(idx), dev(ptr), n_nb, use_schedule=True)
    p_n = so.CSRPattern(dev(idx), dev(ptr), n_nb, use_schedule=False)
    a = so.seg_weighted_pool(data, w1, None, None, pattern=p_s)
    b = so.seg_weighted_pool(data, w1, None, None, pattern=p_n)
    assert rel_err(host(a), host(b)) <= TOL
    lin = so.seg_weighted_pool(data, 2 * w1 + w2, None, None, pattern=p_s)
    assert rel_err(host(lin), host(2 * a + so.seg_weighted_pool(data, w2, None, None, pattern=p_s))) <= TOL
    again = so.seg_weighted_pool(data, w1, None, None, pattern=p_s)
    assert torch.equal(a, again)  # no atomics: bit-identical reruns
    # sum pooling == weighted pooling with unit weights
    ones = torch.ones_like(w1)
    assert rel_err(host(so.seg_pool(data, None, None, pool_type="sum", pattern=p_s)),
                   host(so.seg_weighted_pool(data, ones, None, None, pattern=p_s))) <= TOL
