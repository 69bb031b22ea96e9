"""Developer tool: time the four gather launches of one ML-10M-shaped step under tuning knobs, in one process.
    python tools/sweep_gather.py "gather_variant=0" "gather_variant=1,gather_grid=24" ...
Each argument is a comma-separated list of NAME=VALUE development options (sg_dev_option: gather_variant,
gather_grid); every option is reset to 0 (shipped behaviour) between configurations.  The forward launches write the
PLAIN fp32 [agg | wsum] operand (the in-kernel-split GEMM's input); results that differ from the first configuration
are reported with their max abs difference (the staged variant sums in a different order: ~1e-7 relative)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import stargcn_b200  # noqa: F401,E402
from stargcn_b200 import _lib  # noqa: E402
from stargcn_b200._lib import check  # noqa: E402
from stargcn_b200.graph import MultiLinkCSR  # noqa: E402
from stargcn_b200.seg_op import _p, _stream  # noqa: E402

KNOBS = ("gather_variant", "gather_grid", "gather_threads")


def main():
    configs = sys.argv[1:] or ["gather_variant=0"]
    wl = bench.load_workload(os.environ.get("SWEEP_WORKLOAD", "ml-10m"))
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    R, D = wl["R"], wl["D"]
    sides = []
    for side, x_nb in (("user", wl["x_item"]), ("item", wl["x_user"])):
        csr = MultiLinkCSR(*wl[side], n_nb=x_nb.shape[0], device=dev).prepare(backward=True)
        x = torch.from_numpy(x_nb).to(dev)
        ld = (R * D + R + 31) // 32 * 32
        agg_hi = torch.empty((csr.n_dst, ld), device=dev)
        agg_lo = None
        gagg = torch.randn((csr.n_dst, R * D), device=dev)
        gx = torch.empty((csr.n_nb, D), device=dev)
        sched, tsched = csr.schedule(), csr.t_schedule()
        part, tpart = sched.partial(1, D, extra_per_row=1), tsched.partial(1, D)
        t_indptr, t_src, t_w = csr.transposed()

        def fwd(csr=csr, x=x, agg_hi=agg_hi, agg_lo=agg_lo, ld=ld, sched=sched, part=part):
            check(lib.sg_multilink_agg_fwd_split(_p(agg_hi), _p(agg_lo), ld, _p(x), _p(csr.support), _p(csr.end_points),
                                                 _p(csr.cat_indptr), R, csr.n_dst, csr.n_nb, csr.nnz, D, _p(sched.buf),
                                                 sched.chunk, _p(part), _stream()), "fwd")

        def bwd(csr=csr, gx=gx, gagg=gagg, t_w=t_w, t_src=t_src, t_indptr=t_indptr, tsched=tsched, tpart=tpart):
            check(lib.sg_multilink_agg_bwd(_p(gx), _p(gagg), _p(t_w), _p(t_src), _p(t_indptr), R, csr.n_dst, csr.n_nb,
                                           csr.nnz, D, 1, _p(tsched.buf), tsched.chunk, _p(tpart), _stream()), "bwd")
        sides.append((side, fwd, bwd, agg_hi, gx))
        if side == "user" and os.environ.get("SWEEP_LAYOUT"):
            # layout experiment: the same gather (a) unsplit into the strided [n_dst, R*D] layout and (b) through
            # seg_weighted_pool over the concatenated CSR into a contiguous [R*n_dst, D] (level-major) buffer
            agg = torch.empty((csr.n_dst, R * D), device=dev)
            wsum = torch.empty((csr.n_dst, R), device=dev)
            flat = torch.empty((R * csr.n_dst, D), device=dev)
            part0 = sched.partial(1, D)

            def fwd_strided(csr=csr, x=x, agg=agg, wsum=wsum, sched=sched, part=part):
                check(lib.sg_multilink_agg_fwd(_p(agg), _p(wsum), _p(x), _p(csr.support), _p(csr.end_points),
                                               _p(csr.cat_indptr), R, csr.n_dst, csr.n_nb, csr.nnz, D, _p(sched.buf),
                                               sched.chunk, _p(part), _stream()), "fwd_strided")

            def fwd_contig(csr=csr, x=x, flat=flat, sched=sched, part0=part0):
                check(lib.sg_weighted_pool_fwd(_p(flat), _p(x), _p(csr.support), _p(csr.end_points), _p(csr.cat_indptr),
                                               1, R * csr.n_dst, csr.n_nb, csr.nnz, D, 1, _p(sched.buf), sched.chunk,
                                               _p(part0), _stream()), "fwd_contig")
            sides.append(("user[unsplit strided | contiguous]", fwd_strided, fwd_contig, agg, flat))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = {}
    print(f"{'config':44s} " + " ".join(f"{s[0][:22]}:{d}" for s in sides for d in ("fwd", "bwd")) + "    sum(ms)")
    for cfg in configs:
        for k in KNOBS:
            _lib.dev_option(k, 0)
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            _lib.dev_option(k, int(v))
        times = []
        for side, fwd, bwd, agg_hi, gx in sides:
            for name, fn, outbuf in (("fwd", fwd, agg_hi), ("bwd", bwd, gx)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                tot = 0.0
                for _ in range(10):
                    flush.zero_()                      # evict the tables' L2 lines between repetitions
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record()
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
                times.append(tot / 10)
                key = (side, name)
                if key not in ref:
                    ref[key] = outbuf.clone()
                elif not torch.equal(ref[key], outbuf):
                    print(f"   !! {cfg}: {side} {name} differs from the first configuration "
                          f"(max abs {float((ref[key] - outbuf).abs().max()):.3e})")
        print(f"{cfg:44s} " + " ".join(f"{t:9.4f}" for t in times) + f"   {sum(times):8.4f}")


if __name__ == "__main__":
    main()
