"""Synthetic bipartite rating graphs of the BASELINE.json shapes (there is no network for the
MovieLens files, so every benchmark and large test runs on graphs of the named SHAPE).

Construction follows SURVEY.md §8(d): distinct (user, item) pairs; item popularity
∝ (rank + c)^-alpha, user activity log-normal, every node has degree >= 1; column ids sorted
inside each row (scipy ``tocsr`` order, mxgraph/datasets.py:116-121); rating levels drawn from
a fixed categorical distribution; support = 1/sqrt(d_row * d_col) on the whole matrix
(GraphSampler/graph_sampler.cpp:408-412); the per-level split keeps the within-row order
(graph_sampler.cpp:300-311).  Seed 1000 is the one the reference's own harnesses use
(seg_ops.cu:20, test_seg_ops.py:466).
"""
import numpy as np

# name -> (users, items, edges per direction in the per-iteration train graph, levels, D)   SURVEY §8 table
SHAPES = {
    "ml-100k-d32": (943, 1682, 62_000, 5, 32),
    "ml-100k": (943, 1682, 62_000, 5, 64),
    "ml-1m": (6040, 3706, 710_000, 5, 64),
    "ml-10m": (69_878, 10_677, 8_000_000, 10, 64),
    "douban": (3000, 3000, 110_000, 5, 64),
}
# ML-10M-like rating histogram for 0.5 … 5.0 (synthetic, stated in DESIGN.md); 5-level sets use 1…5
LEVEL_P10 = np.array([.01, .04, .01, .07, .03, .24, .09, .29, .06, .16])
LEVEL_P5 = np.array([.06, .11, .27, .34, .22])


def _csr(rows, cols, vals, n_rows, n_cols):
    order = np.argsort(rows * np.int64(n_cols) + cols, kind="stable")
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n_rows))])
    return indptr.astype(np.int32), cols.astype(np.int32), vals.astype(np.float32), rows.astype(np.int32)


def _sorted_unique(x):
    """np.unique for int64 keys via sort + neighbour compare (numpy 2.3's hash-based unique is ~8x slower here)."""
    x = np.sort(x)
    if x.size == 0:
        return x
    keep = np.empty(x.size, bool)
    keep[0] = True
    np.not_equal(x[1:], x[:-1], out=keep[1:])
    return x[keep]


def make_bipartite(n_user, n_item, n_edges, n_levels=10, seed=1000, alpha=1.0, c=20.0, sigma=1.0):
    """Returns a dict with both CSR directions ('u2i' rows=users, 'i2u' rows=items): indptr, cols,
    vals (rating level value), support, plus degrees and the level values."""
    rng = np.random.default_rng(seed)
    n_edges = int(min(n_edges, n_user * n_item // 2))
    p_item = 1.0 / (np.arange(n_item) + c) ** alpha
    p_item = rng.permutation(p_item / p_item.sum())
    p_user = rng.lognormal(0.0, sigma, n_user)
    p_user /= p_user.sum()
    cdf_user, cdf_item = np.cumsum(p_user), np.cumsum(p_item)
    keys = np.zeros(0, np.int64)
    # every node gets one guaranteed edge; the rest are drawn from the two marginals and de-duplicated
    base = np.concatenate([np.arange(n_user, dtype=np.int64) * n_item + rng.integers(0, n_item, n_user),
                           rng.integers(0, n_user, n_item).astype(np.int64) * n_item + np.arange(n_item)])
    keys = _sorted_unique(base)
    while keys.size < n_edges:
        need = int((n_edges - keys.size) * 1.3) + 1024
        # inverse-CDF sampling (searchsorted) — the same distribution as rng.choice(p=...) at a tenth of the time
        u = np.minimum(np.searchsorted(cdf_user, rng.random(need), side="right"), n_user - 1).astype(np.int64)
        i = np.minimum(np.searchsorted(cdf_item, rng.random(need), side="right"), n_item - 1).astype(np.int64)
        keys = _sorted_unique(np.concatenate([keys, u * n_item + i]))
    if keys.size > n_edges:  # drop extras, never the guaranteed ones
        extra = np.setdiff1d(keys, base, assume_unique=True)
        drop = rng.choice(extra.size, size=keys.size - n_edges, replace=False)
        keys = np.setdiff1d(keys, extra[drop], assume_unique=True)
    u, i = (keys // n_item).astype(np.int64), (keys % n_item).astype(np.int64)
    levels = (np.arange(n_levels) + 1).astype(np.float32) * (0.5 if n_levels == 10 else 1.0)
    p_lvl = LEVEL_P10 if n_levels == 10 else (LEVEL_P5 if n_levels == 5 else np.full(n_levels, 1.0 / n_levels))
    vals = levels[np.minimum(np.searchsorted(np.cumsum(p_lvl / p_lvl.sum()), rng.random(keys.size), side="right"), n_levels - 1)]
    deg_u = np.bincount(u, minlength=n_user).astype(np.int32)
    deg_i = np.bincount(i, minlength=n_item).astype(np.int32)
    out = dict(n_user=n_user, n_item=n_item, nnz=int(keys.size), levels=levels, deg_user=deg_u, deg_item=deg_i)
    for name, (r, cidx, n_r, n_c, dr, dc) in {"u2i": (u, i, n_user, n_item, deg_u, deg_i),
                                              "i2u": (i, u, n_item, n_user, deg_i, deg_u)}.items():
        indptr, cols, v, rows = _csr(r, cidx, vals, n_r, n_c)
        sup = np.sqrt(np.float32(1.0) / dr[rows].astype(np.float32) / dc[cols].astype(np.float32)).astype(np.float32)
        out[name] = dict(indptr=indptr, cols=cols, vals=v, support=sup, rows=rows)
    return out


def split_by_level(indptr, cols, vals, support, levels):
    """Order-preserving per-rating-level split == multi_link_split_by_value followed by the three
    np.take calls of CSRMat.sample_neighbors (mxgraph/graph.py:725-745).  Returns the
    (end_points_l, indptr_l, support_l, positions_l) lists the aggregator consumes."""
    n = indptr.shape[0] - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    ep_l, ptr_l, sup_l, pos_l = [], [], [], []
    for lv in levels:
        pos = np.flatnonzero(vals == lv).astype(np.int32)
        cnt = np.bincount(rows[pos], minlength=n)
        ptr_l.append(np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32))
        ep_l.append(cols[pos].astype(np.int32))
        sup_l.append(support[pos].astype(np.float32))
        pos_l.append(pos)
    return ep_l, ptr_l, sup_l, pos_l


def make_layer_inputs(shape="ml-10m", seed=1000, scale_edges=1.0):
    """Everything one HeterGCN layer needs on a graph of the named shape (full neighbourhood,
    every node selected, so local ids == global ids)."""
    n_user, n_item, n_edges, n_levels, D = SHAPES[shape]
    g = make_bipartite(n_user, n_item, int(n_edges * scale_edges), n_levels, seed)
    rng = np.random.default_rng(seed + 1)
    out = dict(graph=g, D=D, R=n_levels, n_user=n_user, n_item=n_item, nnz=g["nnz"])
    out["x_user"] = rng.standard_normal((n_user, D), dtype=np.float32)
    out["x_item"] = rng.standard_normal((n_item, D), dtype=np.float32)
    # user side aggregates item rows; item side aggregates user rows
    out["user"] = split_by_level(g["u2i"]["indptr"], g["u2i"]["cols"], g["u2i"]["vals"], g["u2i"]["support"], g["levels"])
    out["item"] = split_by_level(g["i2u"]["indptr"], g["i2u"]["cols"], g["i2u"]["vals"], g["i2u"]["support"], g["levels"])
    return out
