"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz from the REFERENCE's own code.

Run in the authoring container (needs /root/reference):  python oracle/gen_golden.py

Sources of truth
  npy:  numpy known-answer functions of the reference's test file
        (/root/reference/seg_ops_cuda/mxnet_op/test_seg_ops.py:11-99) — float64 maths
  ref:  the reference authors' CPU loops and GraphSampler bookkeeping compiled unmodified
        into oracle/_ref/*.so — fp32, the exact accumulation order of the CPU operator
Inputs are NOT stored: they are regenerated from the seed by oracle/cases.py.
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, npy_ref, ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    assert npy_ref.available() and ref.available(), "needs /root/reference and a built oracle/_ref"
    npy = npy_ref.load()
    os.makedirs(OUT, exist_ok=True)
    g = {}

    # ---- contiguous (batch, nnz) ops: npy only (reference has no standalone CPU loop for them) ----
    for ci, (b, s, n) in enumerate(cases.CONTIG_SHAPES):
        c = cases.contig_case(100 + ci, b, s, n)
        d64 = c["data"].astype(np.float64)
        g[f"seg_sum/{ci}/npy"] = npy.npy_seg_sum(d64, c["indptr"])
        g[f"seg_broadcast_add/{ci}/npy"] = npy.npy_seg_broadcast_add(d64, c["rhs"].astype(np.float64), c["indptr"]).astype(np.float32)
        g[f"seg_broadcast_mul/{ci}/npy"] = npy.npy_seg_broadcast_mul(d64, c["rhs"].astype(np.float64), c["indptr"]).astype(np.float32)
        g[f"seg_broadcast_to/{ci}/npy"] = npy.npy_seg_broadcast_to(c["rhs"], c["indptr"], n)
        g[f"seg_softmax/{ci}/npy"] = npy.npy_seg_softmax(d64, c["indptr"])

    # ---- gather ops on the reference's shapes: npy + ref ----
    for ci, shp in enumerate(cases.GATHER_SHAPES):
        b, s, t, n, f = shp
        c = cases.gather_case(200 + ci, *shp, scale=1.0)
        d64, w64 = c["data"].astype(np.float64), c["weights"].astype(np.float64)
        g[f"weighted_pool/{ci}/npy"] = npy.npy_seg_weighted_pool(d64, w64, c["indices"], c["indptr"])
        g[f"weighted_pool/{ci}/ref"] = ref.weighted_pool_fwd(c["data"], c["weights"], c["indices"], c["indptr"])
        g[f"weighted_pool_bwd_data/{ci}/ref"] = ref.weighted_pool_bwd_data(c["gout"], c["weights"], c["indices"], c["indptr"], t)
        kc_npy = npy.npy_seg_take_k_corr(c["embed1"].astype(np.float64), d64, c["indices"], c["indptr"])
        kc_ref = ref.take_k_corr(c["embed1"], c["data"], c["indices"], c["indptr"])
        g[f"take_k_corr/{ci}/npy"] = kc_npy.astype(np.float32)
        g[f"take_k_corr/{ci}/ref"] = kc_ref
        for pt in ("sum", "avg", "max"):
            g[f"seg_pool_{pt}/{ci}/npy"] = npy.npy_seg_pool(d64, c["indices"], c["indptr"], pt)
            val, am = ref.seg_pool_fwd(c["data"], c["indices"], c["indptr"], pt)
            g[f"seg_pool_{pt}/{ci}/ref"] = val
            if pt == "max":
                g[f"seg_pool_max_argmax/{ci}/ref"] = am
            g[f"seg_pool_{pt}_bwd/{ci}/ref"] = ref.seg_pool_bwd(c["gout"], am, c["indices"], c["indptr"], t, pt)
        if n <= 500:  # the reference's explicit max-pool gradient (test_seg_ops.py:87-99)
            c10 = cases.gather_case(200 + ci, *shp, scale=10.0)
            g[f"seg_pool_max_grad/{ci}/npy"] = npy.grad_seg_max_pool(c10["gout"], c10["data"], c10["indices"], c10["indptr"])

    # ---- extra shapes incl. empty segments: ref only (npy mean of an empty slice is NaN) ----
    for ci, shp in enumerate(cases.EXTRA_GATHER_SHAPES):
        b, s, t, n, f = shp
        c = cases.gather_case(300 + ci, *shp, allow_empty=True)
        g[f"x_weighted_pool/{ci}/ref"] = ref.weighted_pool_fwd(c["data"], c["weights"], c["indices"], c["indptr"])
        g[f"x_weighted_pool_bwd_data/{ci}/ref"] = ref.weighted_pool_bwd_data(c["gout"], c["weights"], c["indices"], c["indptr"], t)
        g[f"x_take_k_corr/{ci}/ref"] = ref.take_k_corr(c["embed1"], c["data"], c["indices"], c["indptr"])
        for pt in ("sum", "avg"):   # prototype max leaves -FLT_MAX on empty segments (seg_ops.cu:1058); MXNet op writes 0
            val, am = ref.seg_pool_fwd(c["data"], c["indices"], c["indptr"], pt)
            g[f"x_seg_pool_{pt}/{ci}/ref"] = val
            g[f"x_seg_pool_{pt}_bwd/{ci}/ref"] = ref.seg_pool_bwd(c["gout"], am, c["indices"], c["indptr"], t, pt)

    # ---- GraphSampler integer bookkeeping: ref (bit-exact targets) ----
    for ci, (nr, nc, nnz, nv) in enumerate(cases.GRAPH_SHAPES):
        c = cases.graph_case(400 + ci, nr, nc, nnz, nv)
        g[f"row_indices/{ci}/ref"] = ref.gen_row_indices_by_indptr(c["indptr"], nnz)
        for symm in (0, 1):
            g[f"support_symm{symm}/{ci}/ref"] = ref.get_support(c["row_deg"], c["col_deg"], c["indptr"], c["end_points"], symm)
        idx_l, ptr_l = ref.multi_link_split_by_value(c["values"], c["indptr"], c["levels"])
        g[f"split_indices/{ci}/ref"] = np.concatenate(idx_l).astype(np.int32)
        g[f"split_indptrs/{ci}/ref"] = np.stack(ptr_l).astype(np.int32)
        samp, sptr = ref.random_sample_fix_neighbor(7, c["indptr"], c["sel"], -1)
        g[f"sample_full/{ci}/ref_idx"] = samp
        g[f"sample_full/{ci}/ref_ptr"] = sptr
        ep, val, ptr = ref.remove_edges(c["end_points"], c["values"], c["indptr"], c["rm_rows"], c["rm_cols"])
        g[f"remove_edges/{ci}/ref_ep"] = ep
        g[f"remove_edges/{ci}/ref_val"] = val
        g[f"remove_edges/{ci}/ref_ptr"] = ptr

    path = os.path.join(OUT, "seg_ops_golden.npz")
    np.savez_compressed(path, **{k.replace("/", "__"): cases.sub(v) for k, v in g.items()})
    print(f"wrote {path}: {len(g)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        main()
