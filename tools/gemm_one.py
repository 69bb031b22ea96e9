"""Developer tool: run ONE forward-transform GEMM [69878,650].[250,650]^T a few times (ncu target)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import stargcn_b200  # noqa: F401,E402
from stargcn_b200 import _lib  # noqa: E402
from stargcn_b200._lib import check  # noqa: E402

lib = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    _lib.dev_option(k, int(v))
n, Kx, U, ld = 69_878, 650, 250, 672
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((n, ld), device="cuda", generator=g)
w = torch.randn((U, ld), device="cuda", generator=g)
mask = lambda x: (x.view(torch.int32) & -8192).view(torch.float32)
a_hi, w_hi = mask(a), mask(w)
a_lo, w_lo = a - a_hi, w - w_hi
out = torch.empty((n, U), device="cuda")
for _ in range(4):
    check(lib.sg_gemm_tf32x3(p(out), U, p(a_hi), p(a_lo), ld, p(w_hi), p(w_lo), ld, n, U, Kx, 0, 0, ctypes.c_float(0.0), None, 1,
                             None, st()), "gemm")
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8)()
check(lib.sg_gemm_trace_read(buf), "trace")
names = ["mma:wait full", "mma:wait tmem_empty", "mma:total", "producer0:wait empty", "producer0:total", "epilogue0:wait tmem_full",
         "epilogue0:total", "k-blocks"]
print({n: int(v) // 4 for n, v in zip(names, buf)}, "(cycles per launch, pair 0 leader; 4 launches averaged)")
print("done")
