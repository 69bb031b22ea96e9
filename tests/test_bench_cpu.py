"""bench.py's two CPU arms compute the same function: the reference operator order (FullyConnected +
seg_weighted_pool at F=250, serial data-gradient scatter) and the aggregate-first, row-parallel variant that
is reported beside it (`cpu_baseline.parallel_variant`).  Checked on one rating level of the ML-100k shape."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import segops as orc  # noqa: E402


def test_parallel_cpu_variant_matches_reference_order():
    wl = bench.load_workload("ml-100k")
    R, D = wl["R"], wl["D"]
    ws, bs = bench.make_params(R, D, bench.AGG_UNITS)
    rs = np.random.RandomState(3)
    bs = [rs.uniform(-0.1, 0.1, b.shape).astype(np.float32) for b in bs]      # non-zero bias: exercises wsum
    _kind, pool_fwd, pool_bwd = bench.cpu_pool_functions()
    for side, x_nb, n_dst in (("user", wl["x_item"], wl["n_user"]), ("item", wl["x_user"], wl["n_item"])):
        ep_l, ptr_l, sup_l = wl[side]
        r = 2
        nnz = int(ptr_l[r][-1])
        ep, sup, ptr = np.ascontiguousarray(ep_l[r][:nnz]), np.ascontiguousarray(sup_l[r][:nnz]), ptr_l[r]
        gout = rs.standard_normal((n_dst, bench.AGG_UNITS)).astype(np.float32)
        out_ref, gx_ref = bench.cpu_reference_step(x_nb, ws[r:r + 1], bs[r:r + 1], [ep], [ptr], [sup], gout, pool_fwd, pool_bwd)
        t_indptr, t_perm, t_seg = orc.csr_transpose(ep, ptr, x_nb.shape[0])
        wsum = np.add.reduceat(np.concatenate([sup, [np.float32(0)]]), ptr[:-1].astype(np.int64)).astype(np.float32)
        wsum[np.diff(ptr) == 0] = 0.0
        out_par, gx_par = bench.cpu_parallel_step(x_nb, ws[r], bs[r], ep, ptr, sup, wsum, t_indptr, t_seg,
                                                  np.ascontiguousarray(sup[t_perm]), gout, pool_fwd)
        assert np.abs(out_ref - out_par).max() <= 1e-5 * np.abs(out_ref).max(), side
        assert np.abs(gx_ref - gx_par).max() <= 1e-5 * np.abs(gx_ref).max(), side


def test_reference_arm_line_has_the_contract_keys(capsys, monkeypatch):
    """`bench.py --impl reference` prints one JSON line with the keys the driver reads."""
    import json
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--workload", "ml-100k", "--steps", "1", "--warmup", "0"])
    assert bench.main() == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["parallel_variant"]["value"] > 0
