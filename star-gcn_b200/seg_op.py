"""The ``seg_op`` operator surface of the reference, on torch CUDA tensors.

Mirrors the eight public operators the reference registers as ``mx.nd.contrib.seg_*``
(/root/reference/seg_ops_cuda/README.md:5-175; input-name lists seg_op.cc:361,418,469,515,
566,638,691,787) with the same argument names, shapes, dtypes (float32 data / int32 indices
only) and gradient pairing (FGradient graphs, seg_op.cc:370-379,427-443,478-490,647-659,
700-712,796-813).  Every operator runs a hand-written sm_100a kernel through the C ABI of
include/stargcn_b200.h — there is no eager/CPU fallback.

Beyond the reference surface each operator accepts

  out=, req=      the MXNet write/add/null request on a caller-owned destination
                  (seg_op.cc:188-196); no autograd when ``out`` is given
  pattern=        a :class:`CSRPattern` that caches what depends only on (indices, indptr):
                  the stable transpose used by every backward pass and the load-balancing
                  schedule.  The reference rebuilds this with a radix sort on EVERY backward
                  call (seg_op.cu:882-926); here it is built once per sampled plan.
"""
import ctypes
import weakref
from collections import OrderedDict

import torch

from . import _lib
from ._lib import BCAST, POOL, REDUCE, REQ, check

DEFAULT_CHUNK = 256
_PLAN_MIN_NNZ = 1 << 15  # below this a schedule costs more than it saves


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _chk_float(t, name, ndim):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (no CPU path exists)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype} (seg_op.h:232,410,530)")
    if t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dimensions, got shape {tuple(t.shape)}")
    return t.contiguous()


def _chk_int(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (no CPU path exists)")
    if t.dtype != torch.int32:
        raise TypeError(f"{name} must be int32, got {t.dtype} (seg_op.h:233,413,532)")
    if t.dim() != 1:
        raise ValueError(f"{name} must be 1-D, got shape {tuple(t.shape)}")
    return t.contiguous()


def _bytes(n, device):
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=device)


class Schedule:
    """Device-side work-item schedule for one indptr (sg_plan_build)."""

    def __init__(self, indptr, nnz, chunk=DEFAULT_CHUNK):
        lib = _lib.load()
        self.n_seg = indptr.numel() - 1
        self.nnz = int(nnz)
        self.chunk = int(chunk)
        self.indptr = indptr
        self.nbytes = lib.sg_plan_bytes(self.n_seg, self.nnz, self.chunk)
        self.buf = _bytes(self.nbytes, indptr.device)
        self.partial_rows = int(lib.sg_plan_partial_rows(self.n_seg, self.nnz, self.chunk))
        self.rebuild_()

    def rebuild_(self):
        """(Re)build the schedule from the current contents of ``indptr`` into the same buffer."""
        check(_lib.load().sg_plan_build(_p(self.buf), self.nbytes, _p(self.indptr), self.n_seg, self.nnz, self.chunk,
                                        _stream()), "sg_plan_build")

    def partial(self, K, F, extra_per_row=0):
        """Scratch for the partial rows of split segments.  Taken from the caching allocator on EVERY call, so
        two streams that use the same pattern concurrently never share it (the allocator hands a block to
        another stream only after the kernels that used it have finished)."""
        need = K * self.partial_rows * (F + extra_per_row)
        return torch.empty(max(need, 4), dtype=torch.float32, device=self.buf.device)


class CSRPattern:
    """(indices, indptr) of one CSR neighbour list plus everything derived from it once."""

    def __init__(self, indices, indptr, n_nb, chunk=DEFAULT_CHUNK, use_schedule=None):
        self.indices = _chk_int(indices, "indices")
        self.indptr = _chk_int(indptr, "indptr")
        if self.indptr.numel() < 1:
            raise ValueError("indptr must have at least one element")
        self.n_seg = self.indptr.numel() - 1
        self.nnz = self.indices.numel()
        self.n_nb = int(n_nb)
        self.chunk = int(chunk)
        self.use_schedule = (self.nnz >= _PLAN_MIN_NNZ) if use_schedule is None else bool(use_schedule)
        self._sched = None
        self._t = None
        self._t_sched = None

    # -- forward schedule --
    def schedule(self):
        if not self.use_schedule:
            return None
        if self._sched is None:
            self._sched = Schedule(self.indptr, self.nnz, self.chunk)
        return self._sched

    # -- stable transpose (t_indptr, t_perm, t_seg) --
    def transpose(self):
        if self._t is None:
            lib = _lib.load()
            dev = self.indices.device
            t_indptr = torch.empty(self.n_nb + 1, dtype=torch.int32, device=dev)
            t_perm = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)[: self.nnz]
            t_seg = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)[: self.nnz]
            ws_bytes = lib.sg_csr_transpose_ws_bytes(self.n_seg, self.n_nb, self.nnz)
            if ws_bytes == 0:
                check(2, "sg_csr_transpose_ws_bytes")
            ws = _bytes(ws_bytes, dev)
            check(lib.sg_csr_transpose(_p(t_indptr), _p(t_perm), _p(t_seg), _p(self.indices), _p(self.indptr),
                                       self.n_seg, self.n_nb, self.nnz, _p(ws), ws_bytes, _stream()),
                  "sg_csr_transpose")
            self._t = (t_indptr, t_perm, t_seg)
        return self._t

    def t_schedule(self):
        if not self.use_schedule:
            return None
        if self._t_sched is None:
            self._t_sched = Schedule(self.transpose()[0], self.nnz, self.chunk)
        return self._t_sched


_PATTERN_CACHE = OrderedDict()
_PATTERN_CACHE_MAX = 64
_PATTERN_CACHE_MAX_BYTES = 1 << 30      # derived arrays (transpose + schedules) cost ~40 B per edge


def _pattern_bytes(pat):
    return 40 * pat.nnz + 24 * (pat.n_seg + pat.n_nb)


def get_pattern(indices, indptr, n_nb):
    """Cached CSRPattern for a pair of index tensors (keyed on storage identity + version)."""
    key = (indices.data_ptr(), indptr.data_ptr(), indices._version, indptr._version, indices.numel(),
           indptr.numel(), int(n_nb))
    hit = _PATTERN_CACHE.get(key)
    if hit is not None:
        ref_i, ref_p, pat = hit
        if ref_i() is indices and ref_p() is indptr:
            _PATTERN_CACHE.move_to_end(key)
            return pat
    pat = CSRPattern(indices, indptr, n_nb)
    try:
        _PATTERN_CACHE[key] = (weakref.ref(indices), weakref.ref(indptr), pat)
        while len(_PATTERN_CACHE) > 1 and (len(_PATTERN_CACHE) > _PATTERN_CACHE_MAX or
                                           sum(_pattern_bytes(v[2]) for v in _PATTERN_CACHE.values()) > _PATTERN_CACHE_MAX_BYTES):
            _PATTERN_CACHE.popitem(last=False)
    except TypeError:
        pass
    return pat


def _sched_args(sched, K, F, extra=0):
    if sched is None:
        return ctypes.c_void_p(0), 0, ctypes.c_void_p(0), None
    part = sched.partial(K, F, extra)
    return _p(sched.buf), sched.chunk, _p(part), part


def _dest(out, shape, req, ref):
    if req not in REQ:
        raise ValueError(f"req must be one of {sorted(REQ)}, got {req!r}")
    if out is None:
        if req == "add":
            raise ValueError("req='add' needs an `out` buffer to accumulate into")
        return torch.empty(shape, dtype=torch.float32, device=ref.device)
    if tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 or not out.is_contiguous() or not out.is_cuda:
        raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {tuple(shape)}")
    return out


# --------------------------------------------------------------------------------------------
# raw (non-differentiable) launchers
# --------------------------------------------------------------------------------------------
def _weighted_pool_fwd(data, weights, pat, out=None, req="write"):
    K, n_nb, F = data.shape
    dst = _dest(out, (K, pat.n_seg, F), req, data)
    sched = pat.schedule()
    plan, chunk, part, _keep = _sched_args(sched, K, F)
    check(_lib.load().sg_weighted_pool_fwd(_p(dst), _p(data), _p(weights), _p(pat.indices), _p(pat.indptr), K,
                                           pat.n_seg, n_nb, pat.nnz, F, REQ[req], plan, chunk, part, _stream()),
          "sg_weighted_pool_fwd")
    return dst


def _weighted_pool_bwd_data(gout, weights, pat, n_nb, out=None, req="write"):
    K, n_seg, F = gout.shape
    dst = _dest(out, (K, n_nb, F), req, gout)
    t_indptr, t_perm, t_seg = pat.transpose()
    sched = pat.t_schedule()
    plan, chunk, part, _keep = _sched_args(sched, K, F)
    check(_lib.load().sg_weighted_pool_bwd_data(_p(dst), _p(gout), _p(weights), _p(t_indptr), _p(t_perm), _p(t_seg),
                                                K, n_seg, n_nb, pat.nnz, F, REQ[req], plan, chunk, part, _stream()),
          "sg_weighted_pool_bwd_data")
    return dst


def _take_k_corr(embed1, embed2, pat, out=None, req="write"):
    K, n_node, F = embed1.shape
    dst = _dest(out, (K, pat.nnz), req, embed1)
    check(_lib.load().sg_take_k_corr(_p(dst), _p(embed1), _p(embed2), _p(pat.indices), _p(pat.indptr), K, n_node,
                                     embed2.shape[1], pat.nnz, F, REQ[req], _stream()), "sg_take_k_corr")
    return dst


def _seg_pool_fwd(data, pat, pool_type):
    B, n_nb, F = data.shape
    dst = torch.empty((B, pat.n_seg, F), dtype=torch.float32, device=data.device)
    argmax = torch.empty((B, pat.n_seg, F), dtype=torch.int32, device=data.device) if pool_type == "max" else None
    sched = pat.schedule() if pool_type != "max" else None
    plan, chunk, part, _keep = _sched_args(sched, B, F)
    check(_lib.load().sg_seg_pool_fwd(_p(dst), _p(argmax), _p(data), _p(pat.indices), _p(pat.indptr), B, pat.n_seg,
                                      n_nb, pat.nnz, F, POOL[pool_type], plan, chunk, part, _stream()),
          "sg_seg_pool_fwd")
    return dst, argmax


def _seg_pool_bwd(gout, argmax, pat, n_nb, pool_type, out=None, req="write"):
    B, n_seg, F = gout.shape
    dst = _dest(out, (B, n_nb, F), req, gout)
    t_indptr, t_perm, t_seg = pat.transpose()
    sched = pat.t_schedule() if pool_type != "max" else None
    plan, chunk, part, _keep = _sched_args(sched, B, F)
    check(_lib.load().sg_seg_pool_bwd(_p(dst), _p(gout), _p(argmax), _p(pat.indptr), _p(t_indptr), _p(t_perm),
                                      _p(t_seg), B, n_seg, n_nb, pat.nnz, F, POOL[pool_type], REQ[req], plan, chunk,
                                      part, _stream()), "sg_seg_pool_bwd")
    return dst


def _seg_reduce(data, indptr, kind, out=None, req="write"):
    B, nnz = data.shape
    n_seg = indptr.numel() - 1
    dst = _dest(out, (B, n_seg), req, data)
    check(_lib.load().sg_seg_reduce(_p(dst), _p(data), _p(indptr), B, nnz, n_seg, REDUCE[kind], REQ[req], _stream()),
          "sg_seg_reduce")
    return dst


def _seg_broadcast(lhs, rhs, indptr, op, nnz, out=None, req="write"):
    B, n_seg = rhs.shape
    dst = _dest(out, (B, nnz), req, rhs)
    check(_lib.load().sg_seg_broadcast_binary(_p(dst), _p(lhs), _p(rhs), _p(indptr), B, nnz, n_seg, BCAST[op],
                                              REQ[req], _stream()), "sg_seg_broadcast_binary")
    return dst


def seg_ids(indptr, nnz=None):
    """indptr -> owning segment of every nnz position (GetSegId, seg_op.cu:91-110)."""
    indptr = _chk_int(indptr, "indptr")
    if nnz is None:
        nnz = int(indptr[-1].item())
    out = torch.empty(max(nnz, 1), dtype=torch.int32, device=indptr.device)[:nnz]
    check(_lib.load().sg_seg_ids(_p(out), _p(indptr), indptr.numel() - 1, nnz, _stream()), "sg_seg_ids")
    return out


# --------------------------------------------------------------------------------------------
# autograd pairings (same sub-graphs as the reference's FGradient registrations)
# --------------------------------------------------------------------------------------------
class _SegWeightedPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, weights, pat):
        ctx.pat = pat
        ctx.save_for_backward(data, weights)
        return _weighted_pool_fwd(data, weights, pat)

    @staticmethod
    def backward(ctx, gout):
        data, weights = ctx.saved_tensors
        gout = gout.contiguous()
        gdata = gw = None
        if ctx.needs_input_grad[0]:   # _backward_seg_take_k_corr_embed2 (seg_op.cc:702)
            gdata = _weighted_pool_bwd_data(gout, weights, ctx.pat, data.shape[1])
        if ctx.needs_input_grad[1]:   # seg_take_k_corr(ograd, data, ...) (seg_op.cc:703)
            gw = _take_k_corr(gout, data, ctx.pat)
        return gdata, gw, None


class _SegTakeKCorr(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embed1, embed2, pat):
        ctx.pat = pat
        ctx.save_for_backward(embed1, embed2)
        return _take_k_corr(embed1, embed2, pat)

    @staticmethod
    def backward(ctx, gout):
        embed1, embed2 = ctx.saved_tensors
        gout = gout.contiguous()
        g1 = g2 = None
        if ctx.needs_input_grad[0]:   # seg_weighted_pool(embed2, ograd, ...) (seg_op.cc:649)
            g1 = _weighted_pool_fwd(embed2, gout, ctx.pat)
        if ctx.needs_input_grad[1]:   # _backward_seg_take_k_corr_embed2(ograd, embed1, ...) (seg_op.cc:650)
            g2 = _weighted_pool_bwd_data(embed1, gout, ctx.pat, embed2.shape[1])
        return g1, g2, None


class _SegPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, pat, pool_type):
        out, argmax = _seg_pool_fwd(data, pat, pool_type)
        ctx.pat, ctx.pool_type, ctx.n_nb = pat, pool_type, data.shape[1]
        ctx.argmax = argmax
        return out

    @staticmethod
    def backward(ctx, gout):
        g = _seg_pool_bwd(gout.contiguous(), ctx.argmax, ctx.pat, ctx.n_nb, ctx.pool_type)
        return g, None, None


class _SegSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, indptr):
        ctx.indptr, ctx.nnz = indptr, data.shape[1]
        return _seg_reduce(data, indptr, "sum")

    @staticmethod
    def backward(ctx, gout):  # _backward_seg_sum = broadcast_to (seg_op.cc:372,392)
        return _seg_broadcast(None, gout.contiguous(), ctx.indptr, "to", ctx.nnz), None


class _SegBroadcast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lhs, rhs, indptr, op):
        ctx.indptr, ctx.op = indptr, op
        ctx.save_for_backward(lhs, rhs)
        return _seg_broadcast(lhs, rhs, indptr, op, lhs.shape[1])

    @staticmethod
    def backward(ctx, gout):
        lhs, rhs = ctx.saved_tensors
        gout = gout.contiguous()
        gl = gr = None
        if ctx.op == "add":          # identity / seg_sum (seg_op.cc:427-443)
            if ctx.needs_input_grad[0]:
                gl = gout
            if ctx.needs_input_grad[1]:
                gr = _seg_reduce(gout, ctx.indptr, "sum")
        else:                        # broadcast_mul(og, rhs) / seg_sum(og*lhs) (seg_op.cc:478-490)
            if ctx.needs_input_grad[0]:
                gl = _seg_broadcast(gout, rhs, ctx.indptr, "mul", gout.shape[1])
            if ctx.needs_input_grad[1]:
                gr = _seg_reduce(gout * lhs, ctx.indptr, "sum")
        return gl, gr, None, None


class _SegBroadcastTo(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, indptr, nnz):
        ctx.indptr = indptr
        return _seg_broadcast(None, data, indptr, "to", nnz)

    @staticmethod
    def backward(ctx, gout):         # seg_sum(ograd) (seg_op.cc:524-538)
        return _seg_reduce(gout.contiguous(), ctx.indptr, "sum"), None, None


class _SegSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, indptr):
        B, nnz = data.shape
        out = torch.empty_like(data)
        check(_lib.load().sg_seg_softmax_fwd(_p(out), _p(data), _p(indptr), B, nnz, indptr.numel() - 1, _stream()),
              "sg_seg_softmax_fwd")
        ctx.indptr = indptr
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, gout):         # seg_op.cc:575-584
        (val,) = ctx.saved_tensors
        gout = gout.contiguous()
        B, nnz = val.shape
        g = torch.zeros_like(val)
        check(_lib.load().sg_seg_softmax_bwd(_p(g), _p(gout), _p(val), _p(ctx.indptr), B, nnz, ctx.indptr.numel() - 1,
                                             REQ["write"], _stream()), "sg_seg_softmax_bwd")
        return g, None


# --------------------------------------------------------------------------------------------
# public operators (names / keywords of mx.nd.contrib.seg_*)
# --------------------------------------------------------------------------------------------
def _chk_seg_shapes(indptr, rhs=None, lhs=None):
    """Shape relations the reference's shape inference enforces (seg_op.h:255-319): indptr has n_seg + 1 >= 1
    entries, ``rhs`` is (batch, n_seg) and ``lhs`` (batch, nnz) with the same batch."""
    if indptr.numel() < 1:
        raise ValueError("indptr must have at least one element")
    if rhs is not None and rhs.shape[1] != indptr.numel() - 1:
        raise ValueError(f"per-segment operand has {rhs.shape[1]} columns but indptr describes {indptr.numel() - 1} segments")
    if lhs is not None and rhs is not None and lhs.shape[0] != rhs.shape[0]:
        raise ValueError(f"batch sizes differ: {lhs.shape[0]} vs {rhs.shape[0]}")


def seg_sum(data, indptr, out=None, req="write"):
    """ret[b, i] = sum(data[b, indptr[i]:indptr[i+1]])   (seg_ops_cuda/README.md:5-22)"""
    data, indptr = _chk_float(data, "data", 2), _chk_int(indptr, "indptr")
    _chk_seg_shapes(indptr)
    if out is not None or req != "write":
        return _seg_reduce(data, indptr, "sum", out, req)
    return _SegSum.apply(data, indptr)


def seg_broadcast_add(lhs, rhs, indptr):
    lhs, rhs, indptr = _chk_float(lhs, "lhs", 2), _chk_float(rhs, "rhs", 2), _chk_int(indptr, "indptr")
    _chk_seg_shapes(indptr, rhs, lhs)
    return _SegBroadcast.apply(lhs, rhs, indptr, "add")


def seg_broadcast_mul(lhs, rhs, indptr):
    lhs, rhs, indptr = _chk_float(lhs, "lhs", 2), _chk_float(rhs, "rhs", 2), _chk_int(indptr, "indptr")
    _chk_seg_shapes(indptr, rhs, lhs)
    return _SegBroadcast.apply(lhs, rhs, indptr, "mul")


def seg_broadcast_to(data, indptr, nnz):
    data, indptr = _chk_float(data, "data", 2), _chk_int(indptr, "indptr")
    _chk_seg_shapes(indptr, data)
    if int(nnz) < 0:
        raise ValueError("nnz must be non-negative")
    return _SegBroadcastTo.apply(data, indptr, int(nnz))


def seg_softmax(data, indptr):
    data, indptr = _chk_float(data, "data", 2), _chk_int(indptr, "indptr")
    _chk_seg_shapes(indptr)
    return _SegSoftmax.apply(data, indptr)


def _pattern_of(pattern, indices, indptr, n_nb):
    if pattern is not None:
        return pattern
    return get_pattern(_chk_int(indices, "indices"), _chk_int(indptr, "indptr"), n_nb)


def seg_take_k_corr(embed1, embed2, neighbor_ids, neighbor_indptr, pattern=None, out=None, req="write"):
    """dst[k, j] = <embed1[k, i, :], embed2[k, neighbor_ids[j], :]> for j in segment i."""
    embed1, embed2 = _chk_float(embed1, "embed1", 3), _chk_float(embed2, "embed2", 3)
    if embed1.shape[0] != embed2.shape[0] or embed1.shape[2] != embed2.shape[2]:
        raise ValueError("embed1/embed2 must agree on K and feat_dim (seg_op.h:391-401)")
    pat = _pattern_of(pattern, neighbor_ids, neighbor_indptr, embed2.shape[1])
    if pat.n_seg != embed1.shape[1]:
        raise ValueError("neighbor_indptr must have node_num + 1 entries")
    if out is not None or req != "write":
        return _take_k_corr(embed1, embed2, pat, out, req)
    return _SegTakeKCorr.apply(embed1, embed2, pat)


def seg_weighted_pool(data, weights, indices, indptr, pattern=None, out=None, req="write"):
    """dst[k, i, :] = sum_j weights[k, j] * data[k, indices[j], :] for j in segment i."""
    data, weights = _chk_float(data, "data", 3), _chk_float(weights, "weights", 2)
    pat = _pattern_of(pattern, indices, indptr, data.shape[1])
    if weights.shape != (data.shape[0], pat.nnz):
        raise ValueError(f"weights must have shape (batch, nnz)=({data.shape[0]}, {pat.nnz}) (seg_op.h:437-458)")
    if out is not None or req != "write":
        return _weighted_pool_fwd(data, weights, pat, out, req)
    return _SegWeightedPool.apply(data, weights, pat)


def seg_pool(data, indices, indptr, pool_type="avg", pattern=None):
    """Unweighted segment pooling, pool_type in {'avg', 'sum', 'max'} (seg_op.h:196-209)."""
    if pool_type not in ("avg", "sum", "max"):
        raise ValueError(f"pool_type must be 'avg', 'sum' or 'max', got {pool_type!r}")
    data = _chk_float(data, "data", 3)
    pat = _pattern_of(pattern, indices, indptr, data.shape[1])
    return _SegPool.apply(data, pat, pool_type)


__all__ = ["seg_sum", "seg_broadcast_add", "seg_broadcast_mul", "seg_broadcast_to", "seg_softmax",
           "seg_take_k_corr", "seg_weighted_pool", "seg_pool", "seg_ids", "CSRPattern", "Schedule", "get_pattern"]
