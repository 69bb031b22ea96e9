"""CPU: pins oracle/seg_ops_oracle.c against the reference's own outputs (tests/golden/).

'ref' fixtures come from the reference authors' CPU loops compiled unmodified
(oracle/_ref, seg_ops.cu:788-844,1048-1129) and must match the C restatement BIT-EXACTLY
(same fp32 accumulation order); 'npy' fixtures come from the float64 numpy functions of
the reference's test_seg_ops.py and must match within the reference's own 1e-4 tolerance
(test_seg_ops.py:126,330,398,462).
"""
import numpy as np
import pytest

from oracle import cases, segops
from oracle.cases import sub

NPY_TOL = dict(rtol=1e-4, atol=1e-4)  # the reference's own bar


@pytest.mark.parametrize("ci", range(len(cases.CONTIG_SHAPES)))
def test_contig_ops_vs_npy(golden, ci):
    b, s, n = cases.CONTIG_SHAPES[ci]
    c = cases.contig_case(100 + ci, b, s, n)
    np.testing.assert_allclose(sub(segops.seg_sum(c["data"], c["indptr"])), golden[f"seg_sum/{ci}/npy"], **NPY_TOL)
    np.testing.assert_allclose(sub(segops.seg_broadcast_add(c["data"], c["rhs"], c["indptr"])),
                               golden[f"seg_broadcast_add/{ci}/npy"], **NPY_TOL)
    np.testing.assert_allclose(sub(segops.seg_broadcast_mul(c["data"], c["rhs"], c["indptr"])),
                               golden[f"seg_broadcast_mul/{ci}/npy"], **NPY_TOL)
    np.testing.assert_array_equal(sub(segops.seg_broadcast_to(c["rhs"], c["indptr"], n)),
                                  golden[f"seg_broadcast_to/{ci}/npy"])
    np.testing.assert_allclose(sub(segops.seg_softmax(c["data"], c["indptr"])), golden[f"seg_softmax/{ci}/npy"],
                               **NPY_TOL)


@pytest.mark.parametrize("ci", range(len(cases.GATHER_SHAPES)))
def test_gather_ops_vs_ref_and_npy(golden, ci):
    shp = cases.GATHER_SHAPES[ci]
    b, s, t, n, f = shp
    c = cases.gather_case(200 + ci, *shp)
    out = segops.seg_weighted_pool(c["data"], c["weights"], c["indices"], c["indptr"])
    np.testing.assert_array_equal(sub(out), golden[f"weighted_pool/{ci}/ref"])
    np.testing.assert_allclose(sub(out), golden[f"weighted_pool/{ci}/npy"], **NPY_TOL)
    gd = segops.seg_weighted_pool_bwd_data(c["gout"], c["weights"], c["indices"], c["indptr"], t)
    np.testing.assert_array_equal(sub(gd), golden[f"weighted_pool_bwd_data/{ci}/ref"])
    kc = segops.seg_take_k_corr(c["embed1"], c["data"], c["indices"], c["indptr"])
    np.testing.assert_array_equal(sub(kc), golden[f"take_k_corr/{ci}/ref"])
    np.testing.assert_allclose(sub(kc), golden[f"take_k_corr/{ci}/npy"], **NPY_TOL)
    for pt in ("sum", "avg", "max"):
        val, am = segops.seg_pool(c["data"], c["indices"], c["indptr"], pt, return_argmax=True)
        np.testing.assert_array_equal(sub(val), golden[f"seg_pool_{pt}/{ci}/ref"])
        np.testing.assert_allclose(sub(val), golden[f"seg_pool_{pt}/{ci}/npy"], **NPY_TOL)
        if pt == "max":
            np.testing.assert_array_equal(sub(am), golden[f"seg_pool_max_argmax/{ci}/ref"])
        g = segops.seg_pool_bwd(c["gout"], am, c["indices"], c["indptr"], t, pt)
        np.testing.assert_array_equal(sub(g), golden[f"seg_pool_{pt}_bwd/{ci}/ref"])
    if n <= 500:
        c10 = cases.gather_case(200 + ci, *shp, scale=10.0)
        _, am = segops.seg_pool(c10["data"], c10["indices"], c10["indptr"], "max", return_argmax=True)
        g = segops.seg_pool_bwd(c10["gout"], am, c10["indices"], c10["indptr"], t, "max")
        np.testing.assert_allclose(g, golden[f"seg_pool_max_grad/{ci}/npy"], rtol=2e-3, atol=2e-3)  # :509


@pytest.mark.parametrize("ci", range(len(cases.EXTRA_GATHER_SHAPES)))
def test_ragged_and_empty_segments_vs_ref(golden, ci):
    shp = cases.EXTRA_GATHER_SHAPES[ci]
    b, s, t, n, f = shp
    c = cases.gather_case(300 + ci, *shp, allow_empty=True)
    out = segops.seg_weighted_pool(c["data"], c["weights"], c["indices"], c["indptr"])
    np.testing.assert_array_equal(out, golden[f"x_weighted_pool/{ci}/ref"])
    gd = segops.seg_weighted_pool_bwd_data(c["gout"], c["weights"], c["indices"], c["indptr"], t)
    np.testing.assert_array_equal(gd, golden[f"x_weighted_pool_bwd_data/{ci}/ref"])
    if n > 0:  # the prototype leaves untouched (uninitialised) positions when nnz == 0
        kc = segops.seg_take_k_corr(c["embed1"], c["data"], c["indices"], c["indptr"])
        np.testing.assert_array_equal(kc, golden[f"x_take_k_corr/{ci}/ref"])
    for pt in ("sum", "avg"):
        val, am = segops.seg_pool(c["data"], c["indices"], c["indptr"], pt, return_argmax=True)
        np.testing.assert_array_equal(val, golden[f"x_seg_pool_{pt}/{ci}/ref"])
        g = segops.seg_pool_bwd(c["gout"], None, c["indices"], c["indptr"], t, pt)
        np.testing.assert_array_equal(g, golden[f"x_seg_pool_{pt}_bwd/{ci}/ref"])
    # MXNet-op semantics the prototype does not have: empty max segment -> 0 / -1 (seg_op.cc:264-268)
    val, am = segops.seg_pool(c["data"], c["indices"], c["indptr"], "max", return_argmax=True)
    empty = np.diff(c["indptr"]) == 0
    assert np.all(val[:, empty, :] == 0) and np.all(am[:, empty, :] == -1)
    assert np.all(am[:, ~empty, :] >= 0)


def test_req_semantics():
    """kAddTo accumulates into the existing buffer, kNullOp leaves it alone (seg_op.cc:188-196)."""
    c = cases.gather_case(5, 2, 9, 6, 57, 8, allow_empty=True)
    base = segops.seg_weighted_pool(c["data"], c["weights"], c["indices"], c["indptr"])
    acc = segops.seg_weighted_pool(c["data"], c["weights"], c["indices"], c["indptr"], req="add", init=c["init_out"])
    np.testing.assert_allclose(acc - c["init_out"], base, rtol=1e-5, atol=1e-5)
    gd = segops.seg_weighted_pool_bwd_data(c["gout"], c["weights"], c["indices"], c["indptr"], 6)
    gacc = segops.seg_weighted_pool_bwd_data(c["gout"], c["weights"], c["indices"], c["indptr"], 6, req="add",
                                             init=c["init_data"])
    np.testing.assert_allclose(gacc - c["init_data"], gd, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("ci", range(len(cases.GRAPH_SHAPES)))
def test_bookkeeping_vs_graph_sampler(golden, ci):
    nr, nc, nnz, nv = cases.GRAPH_SHAPES[ci]
    c = cases.graph_case(400 + ci, nr, nc, nnz, nv)
    np.testing.assert_array_equal(segops.seg_ids(c["indptr"]), golden[f"row_indices/{ci}/ref"])
    t_indptr, t_perm, t_seg = segops.csr_transpose(c["end_points"], c["indptr"], nc)
    np.testing.assert_array_equal(np.diff(t_indptr), c["col_deg"])
    np.testing.assert_array_equal(t_seg, c["rows"][t_perm])
    assert np.all(np.diff(c["end_points"][t_perm]) >= 0)
    for n in range(nc):  # stable: ascending original position inside each destination
        assert np.all(np.diff(t_perm[t_indptr[n]:t_indptr[n + 1]]) > 0)
