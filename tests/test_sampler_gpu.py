"""Device-side neighbour sampling / multi-link split / support against the reference's own
host code (oracle/_ref/libgraph_sampler_ref.so = GraphSampler/graph_sampler.cpp compiled unmodified) when it
is present, else against the numpy restatement in this file.  Integer outputs: bit-exact.  Randomised
fan-out (k < degree): structural invariants only — the reference's own draw depends on OpenMP scheduling
(graph_sampler.cpp:765,776)."""
import numpy as np
import pytest
import torch

from oracle import ref

pytestmark = pytest.mark.gpu


def make_graph(seed=0, n_rows=300, n_cols=180, nnz=6000, R=5, hub=True):
    rs = np.random.RandomState(seed)
    flat = np.sort(rs.choice(n_rows * n_cols, size=nnz, replace=False))
    rows, cols = flat // n_cols, (flat % n_cols).astype(np.int32)
    if hub:   # one very long row and a few empty ones
        extra = np.setdiff1d(np.arange(n_cols), cols[rows == 7])
        rows = np.concatenate([rows, np.full(extra.size, 7)])
        cols = np.concatenate([cols, extra.astype(np.int32)])
        keep = ~np.isin(rows, [3, 4, 299])
        rows, cols = rows[keep], cols[keep]
        order = np.lexsort((cols, rows))
        rows, cols = rows[order], cols[order]
    indptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n_rows))]).astype(np.int32)
    levels = (np.arange(R) + 1).astype(np.float32) * 0.5
    vals = levels[rs.randint(0, R, cols.size)]
    return indptr, cols.astype(np.int32), vals, levels, rows.astype(np.int32)


def np_split(edge_values, indptr, possible):
    n = len(indptr) - 1
    rows = np.repeat(np.arange(n), np.diff(indptr))
    idx_l, ptr_l = [], []
    for v in possible:
        pos = np.flatnonzero(edge_values == v).astype(np.int32)
        idx_l.append(pos)
        ptr_l.append(np.concatenate([[0], np.cumsum(np.bincount(rows[pos], minlength=n))]).astype(np.int32))
    return idx_l, ptr_l


def ref_split(edge_values, indptr, possible):
    if ref.available():
        return ref.multi_link_split_by_value(edge_values, indptr, possible)
    return np_split(edge_values, indptr, possible)


def build(indptr, cols, vals, levels, n_cols, rows):
    from stargcn_b200.sampler import DeviceCSR
    rd = np.diff(indptr).astype(np.int32)
    cd = np.bincount(cols, minlength=n_cols).astype(np.int32)
    return DeviceCSR(indptr, cols, vals, levels, n_cols, row_degrees=rd, col_degrees=cd, symm=True), rd, cd


@pytest.mark.parametrize("symm", [True, False])
def test_support_bit_exact(symm):
    indptr, cols, vals, levels, rows = make_graph(1)
    g, rd, cd = build(indptr, cols, vals, levels, 180, rows)
    cd[5] = 0   # a zero degree must give support 0, not inf
    got = g.compute_support(torch.from_numpy(rd).cuda(), torch.from_numpy(cd).cuda(), symm).cpu().numpy()
    if ref.available():
        want = ref.get_support(rd, cd, indptr, cols, symm)
    else:
        r = np.repeat(rd, rd).astype(np.float32)
        with np.errstate(divide="ignore"):
            want = np.sqrt(np.float32(1.0) / r / cd[cols].astype(np.float32)) if symm else np.float32(1.0) / r
        want = np.where((r != 0) & ((cd[cols] != 0) | (not symm)), want, 0).astype(np.float32)
    assert np.array_equal(got.view(np.int32), want.view(np.int32))


@pytest.mark.parametrize("sel_mode", ["all", "subset"])
def test_full_neighbourhood_matches_reference_bit_exact(sel_mode):
    indptr, cols, vals, levels, rows = make_graph(2)
    g, rd, cd = build(indptr, cols, vals, levels, 180, rows)
    rs = np.random.RandomState(3)
    sel = None if sel_mode == "all" else rs.choice(300, 120, replace=False).astype(np.int32)
    sel_np = np.arange(300, dtype=np.int32) if sel is None else sel
    sampled, dst_indptr, n_sel = g.sample_positions(sel, -1, seed=5)
    if ref.available():
        want_s, want_ptr = ref.random_sample_fix_neighbor(5, indptr, sel_np, -1)
    else:
        want_ptr = np.concatenate([[0], np.cumsum(np.diff(indptr)[sel_np])]).astype(np.int32)
        want_s = np.concatenate([np.arange(indptr[r], indptr[r + 1]) for r in sel_np] + [np.zeros(0, int)]).astype(np.int32)
    assert np.array_equal(dst_indptr.cpu().numpy(), want_ptr)
    assert np.array_equal(sampled.cpu().numpy(), want_s)

    cat_indptr, ep_cat, sup_cat, split_index, val_cat = g.split(sampled, dst_indptr, n_sel, want_index=True,
                                                                want_values=True, check_values=True)
    edge_values = vals[want_s]
    idx_l, ptr_l = ref_split(edge_values, want_ptr, levels)
    cat = cat_indptr.cpu().numpy()
    R = len(levels)
    sup_all = g.support.cpu().numpy()
    off = 0
    for r in range(R):
        seg = cat[r * n_sel:(r + 1) * n_sel + 1]
        assert np.array_equal(seg - seg[0], ptr_l[r])                       # per-level ind_ptr
        n_r = len(idx_l[r])
        assert seg[0] == off
        assert np.array_equal(split_index.cpu().numpy()[off:off + n_r], idx_l[r])       # split_indices
        assert np.array_equal(ep_cat.cpu().numpy()[off:off + n_r], cols[want_s][idx_l[r]])       # np.take x 2
        assert np.array_equal(val_cat.cpu().numpy()[off:off + n_r], edge_values[idx_l[r]])
        assert np.array_equal(sup_cat.cpu().numpy()[off:off + n_r], sup_all[want_s][idx_l[r]])
        off += n_r
    assert cat[-1] == len(want_s)


@pytest.mark.parametrize("k", [1, 8, 32, 200])
def test_fixed_fanout_structure_and_determinism(k):
    indptr, cols, vals, levels, rows = make_graph(4)
    g, rd, cd = build(indptr, cols, vals, levels, 180, rows)
    sampled, dst_indptr, n_sel = g.sample_positions(None, k, seed=11)
    s, p = sampled.cpu().numpy(), dst_indptr.cpu().numpy()
    deg = np.diff(indptr)
    assert np.array_equal(np.diff(p), np.minimum(k, deg))                   # counts = min(k, degree)
    for i in range(300):
        row = s[p[i]:p[i + 1]]
        assert np.all((row >= indptr[i]) & (row < indptr[i + 1]))           # inside the row
        assert len(np.unique(row)) == len(row)                              # without replacement
        if deg[i] <= k:
            assert np.array_equal(row, np.arange(indptr[i], indptr[i + 1]))  # full rows in order (cpp:768-771)
    s2 = g.sample_positions(None, k, seed=11)[0].cpu().numpy()
    assert np.array_equal(s, s2)                                            # same seed -> same sample
    s3 = g.sample_positions(None, k, seed=12)[0].cpu().numpy()
    if k < deg.max():
        assert not np.array_equal(s, s3)
    # partition invariance: sampling a subset of rows draws the same neighbours for those rows
    sel = np.array([7, 250, 9, 100], np.int32)
    ss, sp, _ = g.sample_positions(sel, k, seed=11)
    ss, sp = ss.cpu().numpy(), sp.cpu().numpy()
    for j, r in enumerate(sel):
        assert np.array_equal(ss[sp[j]:sp[j + 1]], s[p[r]:p[r + 1]])


def test_uniformity_of_the_draw():
    """Every neighbour of a long row is chosen about equally often over many seeds."""
    indptr, cols, vals, levels, rows = make_graph(6)
    g, rd, cd = build(indptr, cols, vals, levels, 180, rows)
    hits = np.zeros(indptr[8] - indptr[7])
    sel = np.array([7], np.int32)
    n_trials, k = 400, 30
    for seed in range(n_trials):
        s, _, _ = g.sample_positions(sel, k, seed=seed)
        hits[s.cpu().numpy() - indptr[7]] += 1
    expect = n_trials * k / hits.size
    assert hits.min() > 0.5 * expect and hits.max() < 1.6 * expect


def test_sampled_plan_feeds_the_fused_layer():
    """sample_neighbors -> MultiLinkCSR -> aggregator equals the host-built plan of the same sample."""
    from stargcn_b200 import synth
    from stargcn_b200.graph import MultiLinkCSR
    from stargcn_b200.layers import MultiLinkGCNAggregator
    indptr, cols, vals, levels, rows = make_graph(8)
    g, rd, cd = build(indptr, cols, vals, levels, 180, rows)
    csr = g.sample_neighbors(None, 16, seed=3)
    sampled, dst_indptr, n_sel = g.sample_positions(None, 16, seed=3)
    s, p = sampled.cpu().numpy(), dst_indptr.cpu().numpy()
    sup = g.support.cpu().numpy()
    ep_l, ptr_l, sup_l, _ = synth.split_by_level(p, cols[s], vals[s], sup[s], levels)
    host_csr = MultiLinkCSR(ep_l, ptr_l, sup_l, n_nb=180, device="cuda")
    assert torch.equal(csr.cat_indptr, host_csr.cat_indptr)
    assert torch.equal(csr.end_points, host_csr.end_points)
    assert torch.equal(csr.support, host_csr.support)
    agg = MultiLinkGCNAggregator(units=250, num_links=5, act="leaky", ordinal_sharing=False, accum="sum", in_units=64).cuda()
    x = torch.randn(180, 64, device="cuda")
    assert torch.equal(agg(x, csr), agg(x, host_csr))


def test_value_outside_multi_link_is_reported():
    indptr, cols, vals, levels, rows = make_graph(9)
    vals = vals.copy(); vals[10] = 9.75
    g, rd, cd = build(indptr, cols, vals, levels, 180, rows)
    sampled, dst_indptr, n_sel = g.sample_positions(None, -1)
    with pytest.raises(ValueError):
        g.split(sampled, dst_indptr, n_sel, check_values=True)


def test_remove_edges_bit_exact():
    """Per-iteration batch-edge removal (graph_sampler.cpp:154-201) on the device."""
    indptr, cols, vals, levels, rows = make_graph(10)
    g, rd, cd = build(indptr, cols, vals, levels, 180, rows)
    rs = np.random.RandomState(2)
    pick = rs.choice(cols.size, 900, replace=False)
    rm_r = np.concatenate([rows[pick], [5, 7, 299, 3]]).astype(np.int32)       # + pairs that do not exist / empty rows
    rm_c = np.concatenate([cols[pick], [179, 0, 0, 1]]).astype(np.int32)
    new = g.remove_edges(rm_r, rm_c)
    if ref.available():
        want_ep, want_val, want_ptr = ref.remove_edges(cols, vals, indptr, rm_r, rm_c)
    else:
        key = set(zip(rm_r.tolist(), rm_c.tolist()))
        keep = np.array([(r, c) not in key for r, c in zip(rows.tolist(), cols.tolist())])
        want_ep, want_val = cols[keep], vals[keep]
        want_ptr = np.concatenate([[0], np.cumsum(np.bincount(rows[keep], minlength=300))]).astype(np.int32)
    assert np.array_equal(new.ind_ptr.cpu().numpy(), want_ptr)
    assert np.array_equal(new.end_points.cpu().numpy(), want_ep)
    assert np.array_equal(new.values.cpu().numpy(), want_val)
    # support recomputed from the NEW degrees, bit-exact with get_support on the new matrix
    nrd = np.diff(want_ptr).astype(np.int32)
    ncd = np.bincount(want_ep, minlength=180).astype(np.int32)
    if ref.available():
        want_sup = ref.get_support(nrd, ncd, want_ptr, want_ep, True)
        assert np.array_equal(new.support.cpu().numpy().view(np.int32), want_sup.view(np.int32))
    # removing nothing is the identity
    same = g.remove_edges(np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert torch.equal(same.ind_ptr, g.ind_ptr) and torch.equal(same.end_points, g.end_points)
    assert torch.equal(same.support, g.support)


@pytest.mark.parametrize("n,hi", [(1, 5), (50, 8), (5000, 300), (200000, 70000)])
def test_unique_inverse_first_occurrence_order(n, hi):
    """Plan construction primitive (graph_sampler.h:510-534, serial variant): bit-exact."""
    from stargcn_b200.sampler import merge_nodes, unique_inverse
    rs = np.random.RandomState(n)
    data = rs.randint(0, hi, n).astype(np.int32)
    uniq, inv = unique_inverse(torch.from_numpy(data).cuda())
    _, first = np.unique(data, return_index=True)
    want_uniq = data[np.sort(first)]                      # values in order of first appearance
    lut = {int(v): k for k, v in enumerate(want_uniq)}
    want_inv = np.array([lut[int(v)] for v in data], np.int32)
    assert np.array_equal(uniq.cpu().numpy(), want_uniq)
    assert np.array_equal(inv.cpu().numpy(), want_inv)
    assert np.array_equal(uniq.cpu().numpy()[inv.cpu().numpy()], data)
    a, b = torch.from_numpy(data[: n // 2]).cuda(), torch.from_numpy(data[n // 2:]).cuda()
    u2, (ia, ib) = merge_nodes([a, b])
    assert torch.equal(u2, uniq) and torch.equal(torch.cat([ia, ib]), inv)
