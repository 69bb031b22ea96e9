"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol that
include/stargcn_b200.h declares; the host-side pieces that need no device are sane."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "stargcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    import stargcn_b200  # noqa: F401  (imports the package, which loads the library)
    from stargcn_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.lib_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding table out of sync with the header"
    assert _lib.load().sg_abi_version() == 1


def test_size_queries_and_argument_errors_without_gpu():
    from stargcn_b200 import _lib
    lib = _lib.load()
    assert lib.sg_plan_bytes(10, 100, 16) > 0
    assert lib.sg_plan_partial_rows(10, 100, 16) == 2 * (100 // 16) + 1
    assert lib.sg_plan_bytes(-1, 100, 16) == 0
    # invalid arguments are reported through the return code + sg_last_error, never exit()
    rc = lib.sg_weighted_pool_fwd(None, None, None, None, None, 1, 4, 4, 4, 8, 99, None, 0, None, None)
    assert rc == 1 and b"bad req" in lib.sg_last_error()
    rc = lib.sg_seg_pool_fwd(None, None, None, None, None, 1, 4, 4, 4, 8, 7, None, 0, None, None)
    assert rc == 1 and b"pool_type" in lib.sg_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "sg_seg_pool_fwd")


def test_ops_fail_loudly_without_a_device():
    import torch
    from stargcn_b200 import seg_op
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ValueError, match="CUDA"):
        seg_op.seg_sum(torch.zeros(1, 4), torch.tensor([0, 4], dtype=torch.int32))


def test_synthetic_graph_invariants():
    from stargcn_b200 import synth
    g = synth.make_bipartite(300, 200, 6000, n_levels=10, seed=3)
    assert g["nnz"] == 6000 and g["deg_user"].min() >= 1 and g["deg_item"].min() >= 1
    for name, n_rows in (("u2i", 300), ("i2u", 200)):
        c = g[name]
        assert c["indptr"][0] == 0 and c["indptr"][-1] == 6000 and len(c["indptr"]) == n_rows + 1
        for r in range(n_rows):   # column ids strictly increasing inside each row (distinct + sorted)
            assert np.all(np.diff(c["cols"][c["indptr"][r]:c["indptr"][r + 1]]) > 0)
    # both directions hold the same edge set with the same ratings
    u2i = sorted(zip(g["u2i"]["rows"].tolist(), g["u2i"]["cols"].tolist(), g["u2i"]["vals"].tolist()))
    i2u = sorted(zip(g["i2u"]["cols"].tolist(), g["i2u"]["rows"].tolist(), g["i2u"]["vals"].tolist()))
    assert u2i == i2u
    ep_l, ptr_l, sup_l, pos_l = synth.split_by_level(g["u2i"]["indptr"], g["u2i"]["cols"], g["u2i"]["vals"],
                                                     g["u2i"]["support"], g["levels"])
    assert sum(len(e) for e in ep_l) == 6000
    assert np.array_equal(sum(np.diff(p) for p in ptr_l), np.diff(g["u2i"]["indptr"]))


def test_split_and_support_match_graph_sampler_golden(golden):
    """synth.split_by_level / the support formula against the reference GraphSampler outputs."""
    from oracle import cases
    from stargcn_b200 import synth
    for ci, (nr, nc, nnz, nv) in enumerate(cases.GRAPH_SHAPES):
        c = cases.graph_case(400 + ci, nr, nc, nnz, nv)
        _, ptr_l, _, pos_l = synth.split_by_level(c["indptr"], c["end_points"], c["values"],
                                                  np.ones(nnz, np.float32), c["levels"])
        np.testing.assert_array_equal(np.concatenate(pos_l), golden[f"split_indices/{ci}/ref"])
        np.testing.assert_array_equal(np.stack(ptr_l), golden[f"split_indptrs/{ci}/ref"])
        sup = np.sqrt(np.float32(1.0) / c["row_deg"][c["rows"]].astype(np.float32) /
                      c["col_deg"][c["end_points"]].astype(np.float32)).astype(np.float32)
        np.testing.assert_array_equal(sup, golden[f"support_symm1/{ci}/ref"])


def test_layer_oracle_backward_is_the_gradient_of_its_forward():
    """Central finite differences (the reference's own method, test_seg_ops.py:101-114) on the fp64 oracle."""
    from oracle import layers as orl
    rs = np.random.RandomState(0)
    R, D, U, n_dst, n_nb, nnz = 3, 5, 6, 7, 6, 25
    ptr_l, ep_l, sup_l = [], [], []
    for r in range(R):
        cuts = np.sort(rs.randint(0, nnz + 1, n_dst - 1))
        ptr_l.append(np.concatenate([[0], cuts, [nnz]]).astype(np.int32))
        ep_l.append(rs.randint(0, n_nb, nnz).astype(np.int32))
        sup_l.append(rs.uniform(0.1, 1, nnz).astype(np.float32))
    for accum, ordinal in (("sum", False), ("stack", True)):
        Ur = U // R if accum == "stack" else U
        ws = [rs.normal(size=(Ur, D)) for _ in range(R)]
        bs = [rs.normal(size=(Ur,)) for _ in range(R)]
        x = rs.normal(size=(n_nb, D))
        gout = rs.normal(size=(n_dst, U if accum == "sum" else Ur * R))
        f = lambda: float((orl.multilink_aggregator_forward(x, ws, bs, ep_l, ptr_l, sup_l, accum, "leaky", ordinal, fp64=True)[0] * gout).sum())
        _, pre = orl.multilink_aggregator_forward(x, ws, bs, ep_l, ptr_l, sup_l, accum, "leaky", ordinal, fp64=True)
        gx, gws, gbs = orl.multilink_aggregator_backward(x, ws, bs, ep_l, ptr_l, sup_l, gout, pre, accum, "leaky", ordinal, fp64=True)
        eps = 1e-6
        for arr, grad in ((x, gx), (ws[0], gws[0]), (ws[2], gws[2]), (bs[1], gbs[1])):
            for _ in range(6):
                i = tuple(rs.randint(0, s) for s in arr.shape)
                old = arr[i]
                arr[i] = old + eps; fp = f()
                arr[i] = old - eps; fm = f()
                arr[i] = old
                assert abs((fp - fm) / (2 * eps) - grad[i]) < 1e-5
