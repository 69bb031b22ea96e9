"""TEST INFRASTRUCTURE ONLY — seeded input generators shared by gen_golden.py and tests/.

Fixtures in tests/golden/ store only (case name → expected outputs); the inputs are
regenerated from the seed with numpy's frozen legacy ``RandomState`` streams, which keeps
the committed files small.  Shapes are the ones the reference's own tests use
(/root/reference/seg_ops_cuda/mxnet_op/test_seg_ops.py:118,314-316,382-384,450-452) plus
ragged/empty-segment cases the reference's ``rand_indptr`` never produces.
"""
import numpy as np

# (batch, seg_num, nnz)                                  test_seg_ops.py:118,165,225,269
CONTIG_SHAPES = [(1, 5, 10), (10, 50, 100), (4, 1000, 10000)]
# (K|batch, seg_num, total_ind_num, nnz, feat_dim)       test_seg_ops.py:314-316,382-384,450-452
GATHER_SHAPES = [(1, 5, 10, 30, 128), (10, 50, 20, 500, 4), (4, 1000, 10000, 50000, 4)]
# extra: hot-path feature widths, odd widths, long and empty segments
EXTRA_GATHER_SHAPES = [(1, 64, 40, 700, 64), (2, 33, 17, 300, 250), (1, 40, 25, 2500, 32), (3, 9, 6, 57, 7),
                       (1, 12, 5, 0, 16)]


def rand_indptr(rs, seg_num, nnz, allow_empty=False):
    """Reference flavour (test_seg_ops.py:6-9): sorted distinct cut points → no empty segment.
    allow_empty: cut points drawn with replacement, so empty (and leading/trailing empty)
    segments appear, as they do for isolated / cold-start nodes (graph.py:221-222)."""
    if nnz == 0:
        return np.zeros(seg_num + 1, np.int32)
    if allow_empty:
        cuts = np.sort(rs.randint(0, nnz + 1, size=seg_num - 1))
    else:
        cuts = np.sort(rs.choice(np.arange(1, nnz), seg_num - 1, replace=False))
    return np.concatenate([[0], cuts, [nnz]]).astype(np.int32)


def contig_case(seed, batch, seg_num, nnz, allow_empty=False):
    rs = np.random.RandomState(seed)
    return dict(
        data=rs.normal(0, 1, (batch, nnz)).astype(np.float32),
        rhs=rs.normal(0, 1, (batch, seg_num)).astype(np.float32),
        ograd=rs.normal(0, 1, (batch, nnz)).astype(np.float32),
        indptr=rand_indptr(rs, seg_num, nnz, allow_empty),
    )


def gather_case(seed, batch, seg_num, total, nnz, feat, allow_empty=False, scale=1.0):
    rs = np.random.RandomState(seed)
    return dict(
        data=(rs.normal(0, scale, (batch, total, feat))).astype(np.float32),
        embed1=rs.normal(0, 1, (batch, seg_num, feat)).astype(np.float32),
        weights=rs.normal(0, 1, (batch, nnz)).astype(np.float32),
        indices=rs.randint(0, total, size=(nnz,)).astype(np.int32),
        indptr=rand_indptr(rs, seg_num, nnz, allow_empty),
        gout=rs.normal(0, 1, (batch, seg_num, feat)).astype(np.float32),
        init_data=rs.normal(0, 1, (batch, total, feat)).astype(np.float32),
        init_out=rs.normal(0, 1, (batch, seg_num, feat)).astype(np.float32),
    )


def graph_case(seed, n_row, n_col, nnz, n_val=5):
    """A random CSR rating matrix with distinct, sorted column ids per row."""
    rs = np.random.RandomState(seed)
    flat = np.sort(rs.choice(n_row * n_col, size=nnz, replace=False))
    rows, cols = (flat // n_col).astype(np.int32), (flat % n_col).astype(np.int32)
    indptr = np.zeros(n_row + 1, np.int32)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    levels = (np.arange(n_val) + 1).astype(np.float32) * (0.5 if n_val == 10 else 1.0)
    values = levels[rs.randint(0, n_val, size=nnz)]
    n_rm = max(1, nnz // 10)
    rm = rs.choice(nnz, size=n_rm, replace=False)
    return dict(rows=rows, end_points=cols, indptr=indptr, values=values.astype(np.float32), levels=levels,
                row_deg=np.diff(indptr).astype(np.int32),
                col_deg=np.bincount(cols, minlength=n_col).astype(np.int32),
                rm_rows=rows[rm].copy(), rm_cols=cols[rm].copy(),
                sel=rs.permutation(n_row)[: max(1, n_row // 2)].astype(np.int32))


GRAPH_SHAPES = [(8, 6, 20, 5), (60, 40, 700, 5), (300, 200, 12000, 10)]  # last one takes the _omp split path


def sub(a):
    """Fixtures keep every 8th slice along axis 1 of large expected outputs (file size);
    tests apply the same view to what they computed."""
    a = np.asarray(a)
    return a[:, ::8].copy() if (a.ndim >= 2 and a.size > 40000) else a
