"""GPU parity of the tcgen05 3xTF32 GEMM (csrc/gemm.cu) through the C ABI.

Oracle: numpy fp64 matmul of the same fp32 inputs (the exact answer) and the fp32 numpy result the
reference's FullyConnected would give; bar 1e-5 max-normalised (north_star), and the 3xTF32 result
must not be further from the fp64 answer than a plain fp32 GEMM is, plus that tolerance."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def split(x, ld, transpose=False):
    from stargcn_b200 import _lib
    lib = _lib.load()
    rows, cols = x.shape
    orows = cols if transpose else rows
    hi = torch.empty((orows, ld), dtype=torch.float32, device=x.device)
    lo = torch.empty_like(hi)
    _lib.check(lib.sg_split_tf32(_p(hi), _p(lo), ld, _p(x), rows, cols, x.stride(0), int(transpose), _stream()), "sg_split_tf32")
    return hi, lo


def _po(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm(a_hi, a_lo, b_hi, b_lo, M, N, K, mn_major=False, epilogue=0, slope=0.0, splits=1, ldd=None):
    """a_lo / b_lo None = that operand is plain fp32 and is split inside the kernel (tf32x3_gemm_split_kernel)."""
    from stargcn_b200 import _lib
    lib = _lib.load()
    ldd = N if ldd is None else ldd
    D = torch.full((M, ldd), 7.0, dtype=torch.float32, device=a_hi.device)
    ws = None
    if splits > 1:
        ws = torch.empty(lib.sg_gemm_split_ws_bytes(M, N, splits) // 4, dtype=torch.float32, device=a_hi.device)
    _lib.check(lib.sg_gemm_tf32x3(_p(D), ldd, _p(a_hi), _po(a_lo), a_hi.stride(0), _p(b_hi), _po(b_lo), b_hi.stride(0),
                                  M, N, K, int(mn_major), epilogue, ctypes.c_float(slope), None, splits,
                                  _p(ws) if ws is not None else None, _stream()), "sg_gemm_tf32x3")
    return D


def test_split_is_exact():
    x = torch.randn(37, 53, device="cuda") * 3
    hi, lo = split(x, 56)
    assert torch.equal((hi + lo)[:, :53], x)
    assert torch.equal(hi[:, 53:], torch.zeros_like(hi[:, 53:])) and torch.equal(lo[:, 53:], torch.zeros_like(lo[:, 53:]))
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0      # hi is exactly representable in TF32
    hiT, loT = split(x, 40, transpose=True)
    assert torch.equal((hiT + loT)[:, :37], x.t())


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 256, 64), (300, 250, 650), (1000, 640, 250), (70, 33, 20),
                                   (4099, 250, 650)])
def test_kmajor_gemm_matches_fp64(M, N, K):
    rs = np.random.RandomState(M + N + K)
    a = rs.normal(size=(M, K)).astype(np.float32)
    b = rs.normal(size=(N, K)).astype(np.float32)
    lda = (K + 3) // 4 * 4
    a_hi, a_lo = split(torch.from_numpy(a).cuda(), lda)
    b_hi, b_lo = split(torch.from_numpy(b).cuda(), lda)
    D = gemm(a_hi, a_lo, b_hi, b_lo, M, N, K).cpu().numpy()
    ref64 = a.astype(np.float64) @ b.astype(np.float64).T
    ref32 = a @ b.T
    assert rel_err(D, ref64) <= TOL
    assert rel_err(D, ref32) <= TOL
    assert rel_err(D, ref64) <= rel_err(ref32, ref64) + 2e-6


def test_kmajor_leaky_epilogue_and_padded_output():
    M, N, K = 515, 250, 650
    rs = np.random.RandomState(3)
    a = rs.normal(size=(M, K)).astype(np.float32)
    b = rs.normal(size=(N, K)).astype(np.float32)
    a_hi, a_lo = split(torch.from_numpy(a).cuda(), 672)
    b_hi, b_lo = split(torch.from_numpy(b).cuda(), 672)
    D = gemm(a_hi, a_lo, b_hi, b_lo, M, N, K, epilogue=1, slope=0.1, ldd=256).cpu().numpy()
    z = a.astype(np.float64) @ b.astype(np.float64).T
    ref = np.where(z > 0, z, 0.1 * z)
    assert rel_err(D[:, :N], ref) <= TOL
    assert np.all(D[:, N:] == 7.0)          # padding columns of the destination are untouched


@pytest.mark.parametrize("Kdim,M,N,splits", [(64, 128, 256, 1), (1000, 250, 650, 1), (5000, 250, 650, 7), (333, 40, 70, 3)])
def test_mnmajor_splitk_gemm(Kdim, M, N, splits):
    """D[M,N] = A[K,M]^T . B[K,N] — the weight-gradient shape (reduction over the node axis)."""
    rs = np.random.RandomState(Kdim + M)
    a = rs.normal(size=(Kdim, M)).astype(np.float32)
    b = rs.normal(size=(Kdim, N)).astype(np.float32)
    lda, ldb = (M + 31) // 32 * 32, (N + 31) // 32 * 32
    a_hi, a_lo = split(torch.from_numpy(a).cuda(), lda)
    b_hi, b_lo = split(torch.from_numpy(b).cuda(), ldb)
    D = gemm(a_hi, a_lo, b_hi, b_lo, M, N, Kdim, mn_major=True, splits=splits).cpu().numpy()
    ref64 = a.astype(np.float64).T @ b.astype(np.float64)
    assert rel_err(D, ref64) <= TOL
    D2 = gemm(a_hi, a_lo, b_hi, b_lo, M, N, Kdim, mn_major=True, splits=splits).cpu().numpy()
    assert np.array_equal(D, D2)            # fixed-order split-K reduction: bit-identical reruns


def test_act_bwd_split():
    from stargcn_b200 import _lib
    lib = _lib.load()
    M, U, ldz = 77, 250, 256
    out = torch.randn(M, U, device="cuda")
    gout = torch.randn(M, U, device="cuda")
    hi = torch.empty(M, ldz, device="cuda")
    lo = torch.empty(M, ldz, device="cuda")
    _lib.check(lib.sg_act_bwd_split(_p(hi), _p(lo), ldz, _p(gout), _p(out), M, U, ctypes.c_float(0.1), _stream()), "sg_act_bwd_split")
    ref = torch.where(out > 0, gout, 0.1 * gout)
    assert torch.equal((hi + lo)[:, :U], ref)
    assert float((hi + lo)[:, U:].abs().max()) == 0.0


def _padded(x, ld):
    """x in a buffer whose rows are ld floats apart; the padding holds NaN (it must never be read as data)."""
    buf = torch.full((x.shape[0], ld), float("nan"), dtype=torch.float32, device="cuda")
    buf[:, :x.shape[1]] = torch.from_numpy(x).cuda()
    return buf


@pytest.mark.parametrize("raw_b", [False, True])
@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (300, 250, 650), (1000, 640, 250), (70, 33, 20), (4099, 250, 650),
                                   (69_878, 250, 650)])
def test_kmajor_gemm_raw_operands_split_in_kernel(M, N, K, raw_b):
    """A (and optionally B) as plain fp32: the hi/lo split happens in shared memory.  Same bars as the pre-split
    kernel, and the result is bit-identical to it (same products, same accumulation order)."""
    rs = np.random.RandomState(M + N + K + 1)
    a = rs.normal(size=(M, K)).astype(np.float32)
    b = rs.normal(size=(N, K)).astype(np.float32)
    lda = (K + 3) // 4 * 4 + 4
    a_raw, b_raw = _padded(a, lda), _padded(b, lda)
    b_hi, b_lo = split(torch.from_numpy(b).cuda(), lda)
    if raw_b:
        D = gemm(a_raw, None, b_raw, None, M, N, K)
    else:
        D = gemm(a_raw, None, b_hi, b_lo, M, N, K)
    a_hi, a_lo = split(torch.from_numpy(a).cuda(), lda)
    D_pre = gemm(a_hi, a_lo, b_hi, b_lo, M, N, K)
    Dh = D.cpu().numpy()
    ref64 = a.astype(np.float64) @ b.astype(np.float64).T
    ref32 = a @ b.T
    assert np.isfinite(Dh).all()
    assert rel_err(Dh, ref64) <= TOL
    assert rel_err(Dh, ref64) <= rel_err(ref32, ref64) + 2e-6
    assert torch.equal(D, D_pre)


def test_kmajor_raw_leaky_bias_free_epilogue():
    M, N, K = 515, 250, 650
    rs = np.random.RandomState(4)
    a = rs.normal(size=(M, K)).astype(np.float32)
    b = rs.normal(size=(N, K)).astype(np.float32)
    b_hi, b_lo = split(torch.from_numpy(b).cuda(), 672)
    D = gemm(_padded(a, 672), None, b_hi, b_lo, M, N, K, epilogue=1, slope=0.1, ldd=256).cpu().numpy()
    z = a.astype(np.float64) @ b.astype(np.float64).T
    assert rel_err(D[:, :N], np.where(z > 0, z, 0.1 * z)) <= TOL
    assert np.all(D[:, N:] == 7.0)


@pytest.mark.parametrize("Kdim,M,N,splits", [(64, 128, 256, 1), (1000, 250, 650, 1), (5000, 250, 650, 7), (333, 40, 70, 3),
                                             (69_878, 250, 650, 24)])
def test_mnmajor_gemm_raw_operands(Kdim, M, N, splits):
    """The weight gradient with BOTH operands plain fp32 (gZ and [agg | wsum])."""
    rs = np.random.RandomState(Kdim + M + 1)
    a = rs.normal(size=(Kdim, M)).astype(np.float32)
    b = rs.normal(size=(Kdim, N)).astype(np.float32)
    lda, ldb = (M + 3) // 4 * 4 + 4, (N + 31) // 32 * 32
    D = gemm(_padded(a, lda), None, _padded(b, ldb), None, M, N, Kdim, mn_major=True, splits=splits)
    ref64 = a.astype(np.float64).T @ b.astype(np.float64)
    assert rel_err(D.cpu().numpy(), ref64) <= TOL
    D2 = gemm(_padded(a, lda), None, _padded(b, ldb), None, M, N, Kdim, mn_major=True, splits=splits)
    assert torch.equal(D, D2)
    a_hi, a_lo = split(torch.from_numpy(a).cuda(), lda)
    b_hi, b_lo = split(torch.from_numpy(b).cuda(), ldb)
    assert torch.equal(D, gemm(a_hi, a_lo, b_hi, b_lo, M, N, Kdim, mn_major=True, splits=splits))


def test_act_bwd_unsplit():
    from stargcn_b200 import _lib
    lib = _lib.load()
    M, U, ldz = 77, 250, 252
    out = torch.randn(M, U, device="cuda")
    gout = torch.randn(M, U, device="cuda")
    gz = torch.empty(M, ldz, device="cuda")
    _lib.check(lib.sg_act_bwd_split(_p(gz), None, ldz, _p(gout), _p(out), M, U, ctypes.c_float(0.1), _stream()), "sg_act_bwd_split")
    assert torch.equal(gz[:, :U], torch.where(out > 0, gout, 0.1 * gout)) and float(gz[:, U:].abs().max()) == 0.0
