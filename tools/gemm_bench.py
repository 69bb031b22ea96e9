"""Developer tool: the three GEMM shapes of the ML-10M user-side layer in isolation, per operand path.
    python tools/gemm_bench.py [reps]
Prints ms per call (CUDA events, L2 flushed between repetitions) for: pre-split operands (shipped path), raw A
(split in the kernel), raw A and B; a check whether the tensor core TRUNCATES raw fp32 inputs to TF32 (then the
raw tile itself can stand in for the 'hi' operand); and the TMA delivery rate per SM (sg_tma_probe) as a function of
ring depth, boxes per barrier phase and the number of issuing warps.  Results of round 2: profiles/r02_summary.md."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import stargcn_b200  # noqa: F401,E402
from stargcn_b200 import _lib  # noqa: E402
from stargcn_b200._lib import check  # noqa: E402


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


lib = _lib.load()


def split(x, ld, transpose=False):
    rows, cols = x.shape
    hi = torch.empty((cols if transpose else rows, ld), device=x.device)
    lo = torch.empty_like(hi)
    check(lib.sg_split_tf32(_p(hi), _p(lo), ld, _p(x), rows, cols, x.stride(0), int(transpose), _stream()), "split")
    return hi, lo


def gemm(D, a_hi, a_lo, b_hi, b_lo, M, N, K, mn=False, splits=1, ws=None):
    check(lib.sg_gemm_tf32x3(_p(D), D.stride(0), _p(a_hi), _p(a_lo), a_hi.stride(0), _p(b_hi), _p(b_lo), b_hi.stride(0),
                             M, N, K, int(mn), 0, ctypes.c_float(0.0), None, splits, _p(ws), _stream()), "gemm")


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    dev = torch.device("cuda", 0)
    n, R, D, U = 69_878, 10, 64, 250
    Kx, ld, ldz = R * D + R, 672, 252
    g = torch.Generator(device=dev).manual_seed(0)
    agg = torch.randn((n, ld), device=dev, generator=g)
    gz = torch.randn((n, ldz), device=dev, generator=g)
    w = torch.randn((U, Kx), device=dev, generator=g) * 0.2
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    agg_hi, agg_lo = split(agg[:, :Kx].contiguous(), ld)
    agg_raw = agg_hi + agg_lo          # zero padding, same values
    gz_hi, gz_lo = split(gz[:, :U].contiguous(), ldz)
    gz_raw = gz_hi + gz_lo
    w_hi, w_lo = split(w, ld)
    wt_hi, wt_lo = split(w[:, :R * D].contiguous(), ldz, transpose=True)
    w_raw, wt_raw = w_hi + w_lo, wt_hi + wt_lo
    out = torch.empty((n, U), device=dev)
    gagg = torch.empty((n, R * D), device=dev)
    gw = torch.empty((U, Kx), device=dev)
    splits = 24
    ws = torch.empty(lib.sg_gemm_split_ws_bytes(U, Kx, splits) // 4, device=dev)

    # does the tensor core truncate raw fp32 to TF32?  hi := raw (unmasked), lo := raw - trunc(raw)
    ref = torch.empty_like(out)
    gemm(ref, agg_hi, agg_lo, w_hi, w_lo, n, U, Kx)
    t1 = torch.empty_like(out)
    gemm(t1, agg_raw, agg_lo, w_hi, w_lo, n, U, Kx)
    t2 = torch.empty_like(out)
    gemm(t2, agg_raw, agg_lo, w_raw, w_lo, n, U, Kx)
    torch.cuda.synchronize()
    print(f"truncation check: raw A as hi -> max|diff| {float((t1 - ref).abs().max()):.3e}; raw A and raw B as hi -> "
          f"{float((t2 - ref).abs().max()):.3e}   (max|ref| {float(ref.abs().max()):.3e}; 0 = hardware truncates)")

    cases = {
        "fwd  [69878,650].[250,650]^T": [
            ("pre-split", lambda: gemm(out, agg_hi, agg_lo, w_hi, w_lo, n, U, Kx)),
            ("raw A", lambda: gemm(out, agg_raw, None, w_hi, w_lo, n, U, Kx)),
            ("raw A+B", lambda: gemm(out, agg_raw, None, w_raw, None, n, U, Kx))],
        "dAgg [69878,250].[640,250]^T": [
            ("pre-split", lambda: gemm(gagg, gz_hi, gz_lo, wt_hi, wt_lo, n, R * D, U)),
            ("raw A", lambda: gemm(gagg, gz_raw, None, wt_hi, wt_lo, n, R * D, U)),
            ("raw A+B", lambda: gemm(gagg, gz_raw, None, wt_raw, None, n, R * D, U))],
        "dW   [69878,250]^T.[69878,650] split-K 24": [
            ("pre-split", lambda: gemm(gw, gz_hi, gz_lo, agg_hi, agg_lo, U, Kx, n, mn=True, splits=splits, ws=ws)),
            ("raw A (B pre-split)", lambda: gemm(gw, gz_raw, None, agg_hi, agg_lo, U, Kx, n, mn=True, splits=splits, ws=ws)),
            ("raw A+B", lambda: gemm(gw, gz_raw, None, agg_raw, None, U, Kx, n, mn=True, splits=splits, ws=ws))],
    }
    # --- what can TMA deliver to an SM?  (ring of `stages` x `boxes` 16-KB boxes, no MMA) ---
    small = agg_hi[:32768]                                        # 88 MB: L2-resident after the first pass
    for src, tag in ((small, "L2-resident 88 MB"), (agg_hi, "188 MB (DRAM)")):
        for stages, boxes, same, prod in ((3, 4, 0, 1), (3, 4, 0, 2), (3, 4, 0, 4), (3, 3, 0, 1), (3, 3, 0, 3), (3, 2, 0, 2), (6, 2, 0, 2)):
            iters = 400
            fn = lambda: check(lib.sg_tma_probe(_p(src), src.shape[0], 640, ld, stages + 100 * (prod - 1), boxes,
                                                -iters if same else iters, 148, _stream()), "tma_probe")
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            by = 148 * iters * boxes * 16384
            print(f"tma probe {tag}, {prod} producer warp(s): {stages} stages x {boxes} boxes ({stages * boxes * 16} KB in flight): "
                  f"{by / ms / 1e6 / 148:.1f} GB/s per SM, {by / ms / 1e9:.2f} TB/s total")
    for name, variants in cases.items():
        print(name + ":  " + "   ".join(f"{v}: {timeit(fn, reps, flush):.4f} ms" for v, fn in variants))


if __name__ == "__main__":
    main()
