"""Device-resident CSR structures that feed the fused aggregation kernels.

``MultiLinkCSR`` is the device form of one ``(src_key, dst_key)`` entry of the reference's
computing plan — the ``[end_points_l, edge_values_l, ind_ptr_l, support_l]`` lists that
``StackedHeterGCNLayers.gen_plan`` emits (mxgraph/layers/layers.py:303-336) and
``heter_sage`` re-uploads with four ``nd.array`` copies per rating level on every call
(layers.py:366-377).  Here the R per-level CSRs are concatenated relation-major ONCE
(segment id = r * n_dst + i), together with the stable transpose and the load-balancing
schedules that the backward pass needs, and stay resident for every forward/backward that
uses the plan.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check
from .seg_op import DEFAULT_CHUNK, Schedule, _bytes, _p, _stream


def _as_tensor(a, dtype):
    """numpy array / torch tensor (any device) -> 1-D torch tensor of ``dtype``; no device round trip."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
        return t if t.dtype == dtype else t.to(dtype)
    np_dtype = {torch.int32: np.int32, torch.float32: np.float32}[dtype]
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np_dtype))


class MultiLinkCSR:
    """R rating-level CSR neighbour lists over the same n_dst destination nodes.

    Parameters mirror the aggregator call (aggregators.py:111-128): ``end_points_l[r]`` are
    LOCAL row ids into the neighbour feature matrix, ``indptr_l[r]`` has n_dst+1 entries,
    ``support_l[r]`` is the edge normalisation 1/sqrt(d_i d_j).  Levels without edges may
    arrive as the reference's length-1 dummies (``empty_as_zero``, graph.py:221-222): only
    the first ``indptr_l[r][-1]`` entries of each list are used.

    The lists may be numpy arrays, (pinned) CPU tensors or CUDA tensors.  Each level is copied ONCE, straight
    into its slice of the concatenated device arrays (asynchronously from pinned memory), and the concatenated
    indptr is assembled on the device — no host-side concatenation and no device->host transfer.  The only
    host-side value needed is the edge count of every level: read from the host ``indptr_l`` for free, or given
    as ``nnz_l`` when the lists already live on the device (otherwise ONE stacked device->host read).
    ``validate=True`` additionally checks the index ranges on the device and raises (that check synchronises).
    """

    def __init__(self, end_points_l, indptr_l, support_l, n_nb, device=None, chunk=DEFAULT_CHUNK,
                 use_schedule=True, nnz_l=None, validate=False):
        if not (len(end_points_l) == len(indptr_l) == len(support_l)) or len(indptr_l) == 0:
            raise ValueError("end_points_l, indptr_l and support_l must be non-empty lists of equal length")
        if device is None:
            device = next((t.device for t in list(end_points_l) + list(indptr_l)
                           if isinstance(t, torch.Tensor) and t.is_cuda), torch.device("cuda"))
        self.device = dev = torch.device(device)
        self.R = R = len(indptr_l)
        ptrs = [_as_tensor(p, torch.int32) for p in indptr_l]
        self.n_dst = n_dst = int(ptrs[0].shape[0]) - 1
        if any(p.dim() != 1 or p.shape[0] != n_dst + 1 for p in ptrs):
            raise ValueError("every indptr must have n_dst + 1 entries")
        if R * n_dst >= 2 ** 31:
            raise ValueError("R * n_dst overflows int32")
        if nnz_l is None:
            if all(not p.is_cuda for p in ptrs):
                nnz_l = [int(p[-1]) for p in ptrs]
            else:   # device-resident lists without counts: one stacked read (pass nnz_l to avoid it)
                nnz_l = [int(v) for v in torch.stack([p[-1].to(dev) for p in ptrs]).tolist()]
        self.nnz_l = [int(n) for n in nnz_l]
        if len(self.nnz_l) != R or any(n < 0 for n in self.nnz_l):
            raise ValueError("nnz_l must hold one non-negative edge count per level")
        self.nnz = nnz = int(sum(self.nnz_l))
        self.n_nb = int(n_nb)
        offs = [0]
        for n in self.nnz_l:
            offs.append(offs[-1] + n)
        ep_cat = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
        sup_cat = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)[:nnz]
        ptr2 = torch.empty((R, n_dst + 1), dtype=torch.int32, device=dev)
        for r in range(R):
            n = self.nnz_l[r]
            e, s_ = _as_tensor(end_points_l[r], torch.int32), _as_tensor(support_l[r], torch.float32)
            if e.shape[0] < n or s_.shape[0] < n:
                raise ValueError(f"level {r}: indptr ends at {n} but only {min(e.shape[0], s_.shape[0])} edges were given")
            if n:
                ep_cat[offs[r]:offs[r + 1]].copy_(e[:n], non_blocking=True)
                sup_cat[offs[r]:offs[r + 1]].copy_(s_[:n], non_blocking=True)
            ptr2[r].copy_(ptrs[r], non_blocking=True)
        # cat_indptr[r * n_dst + i] = offs[r] + indptr_r[i];  the last entry closes the last level
        offs_dev = torch.tensor(offs[:-1], dtype=torch.int32).to(dev, non_blocking=True)
        cat_indptr = torch.empty(R * n_dst + 1, dtype=torch.int32, device=dev)
        cat_indptr[:R * n_dst].view(R, n_dst).copy_(ptr2[:, :n_dst] + offs_dev[:, None])
        cat_indptr[R * n_dst:].fill_(nnz)
        if validate:
            bad_ptr = (ptr2[:, 0] != 0).any() | (ptr2[:, 1:] < ptr2[:, :-1]).any() | \
                (ptr2[:, -1] != torch.tensor(self.nnz_l, dtype=torch.int32).to(dev)).any()
            bad_ep = ((ep_cat < 0) | (ep_cat >= self.n_nb)).any() if nnz else torch.zeros((), dtype=torch.bool, device=dev)
            flags = torch.stack([bad_ptr, bad_ep]).tolist()
            if flags[0]:
                raise ValueError("inconsistent indptr lists")
            if flags[1]:
                raise ValueError("end point index out of range of the neighbour feature matrix")
        self._init_device(ep_cat, sup_cat, cat_indptr, chunk, use_schedule)
        self._ptr2, self._offs_dev, self._offs = ptr2, offs_dev, offs

    ZERO_COPY_MAX_BYTES = 16 << 20

    def load_lists_(self, end_points_l, indptr_l, support_l, zero_copy=None):
        """Refresh the plan IN PLACE from new per-level lists with the SAME per-level edge counts (same-shaped plan
        of the next iteration): every level goes straight into its slice of the existing device arrays and the
        concatenated indptr is re-assembled on the device.  The derived structures are stale afterwards — call
        :meth:`rebuild_` (CUDA-graph capturable) before the next use.

        Transport: asynchronous DMA copies from pinned memory (one per list), or — ``zero_copy`` — ONE kernel that
        reads all 3*R pinned host arrays directly (``sg_upload_segments``).  ``zero_copy=None`` picks the kernel for
        plans below ``ZERO_COPY_MAX_BYTES`` whose lists are all pinned CPU tensors: there the ~60 DMA set-ups, not
        the bytes, are what an upload costs."""
        if getattr(self, "_ptr2", None) is None:
            raise ValueError("load_lists_ needs a plan that was built from per-level lists")
        R, n_dst, offs = self.R, self.n_dst, self._offs
        if not (len(end_points_l) == len(indptr_l) == len(support_l) == R):
            raise ValueError("the refreshed lists must have the same number of levels")
        jobs = []
        for r in range(R):
            n = self.nnz_l[r]
            e, s_, p = (_as_tensor(end_points_l[r], torch.int32), _as_tensor(support_l[r], torch.float32),
                        _as_tensor(indptr_l[r], torch.int32))
            if p.shape[0] != n_dst + 1 or e.shape[0] < n or s_.shape[0] < n:
                raise ValueError(f"level {r}: shape differs from the plan being refreshed")
            if n:
                jobs.append((self.end_points[offs[r]:offs[r + 1]], e[:n]))
                jobs.append((self.support[offs[r]:offs[r + 1]], s_[:n]))
            jobs.append((self._ptr2[r], p))
        pinned = all((not src.is_cuda) and src.is_pinned() and src.is_contiguous() for _, src in jobs)
        if zero_copy is None:
            zero_copy = pinned and sum(src.numel() * 4 for _, src in jobs) <= self.ZERO_COPY_MAX_BYTES
        if zero_copy:
            if not pinned:
                raise ValueError("zero_copy needs contiguous pinned CPU tensors")
            n = len(jobs)
            dsts = (ctypes.c_void_p * n)(*[dst.data_ptr() for dst, _ in jobs])
            srcs = (ctypes.c_void_p * n)(*[src.data_ptr() for _, src in jobs])
            sizes = (ctypes.c_size_t * n)(*[src.numel() * 4 for _, src in jobs])
            check(_lib.load().sg_upload_segments(dsts, srcs, sizes, n, _stream()), "sg_upload_segments")
            self._keep_alive = [src for _, src in jobs]      # the kernel reads them asynchronously
        else:
            for dst, src in jobs:
                dst.copy_(src, non_blocking=True)
        self.cat_indptr[:R * n_dst].view(R, n_dst).copy_(self._ptr2[:, :n_dst] + self._offs_dev[:, None])
        return self

    @classmethod
    def from_device(cls, end_points, support, cat_indptr, R, n_dst, n_nb, chunk=DEFAULT_CHUNK, use_schedule=True):
        """Adopt already-concatenated device arrays (the device sampler's output)."""
        self = cls.__new__(cls)
        self.device = end_points.device
        self.R, self.n_dst, self.n_nb = int(R), int(n_dst), int(n_nb)
        self.nnz = int(end_points.numel())
        self.nnz_l = None
        self._init_device(end_points, support, cat_indptr, chunk, use_schedule)
        return self

    def _init_device(self, end_points, support, cat_indptr, chunk, use_schedule):
        dev = self.device
        self.end_points = end_points.to(dev, torch.int32, non_blocking=True).contiguous()
        self.support = support.to(dev, torch.float32, non_blocking=True).contiguous()
        self.cat_indptr = cat_indptr.to(dev, torch.int32, non_blocking=True).contiguous()
        self.n_seg = self.R * self.n_dst
        self.chunk = int(chunk)
        self.use_schedule = bool(use_schedule)
        self._sched = None
        self._t = None
        self._t_ws = None
        self._t_sched = None
        self.keep_transpose_scratch = False   # True: keep the sort outputs (~8 B/edge) so refresh_weights_ is one launch
        self._ptr2 = None
        self.h2d_bytes = 4 * (2 * self.nnz + self.n_seg + 1)

    def to_lists(self):
        """The reference's per-level ``(end_points_l, indptr_l, support_l)`` lists as numpy arrays (inspection /
        tests: a device -> host copy)."""
        ptr = self.cat_indptr.cpu().numpy().astype(np.int64)
        ep, sup = self.end_points.cpu().numpy(), self.support.cpu().numpy()
        ep_l, ptr_l, sup_l = [], [], []
        for r in range(self.R):
            seg = ptr[r * self.n_dst:(r + 1) * self.n_dst + 1]
            ep_l.append(ep[seg[0]:seg[-1]].copy())
            sup_l.append(sup[seg[0]:seg[-1]].copy())
            ptr_l.append((seg - seg[0]).astype(np.int32))
        return ep_l, ptr_l, sup_l

    def schedule(self):
        if self.use_schedule and self._sched is None:
            self._sched = Schedule(self.cat_indptr, self.nnz, self.chunk)
        return self._sched

    def _build_transposed(self, out=None):
        """(t_indptr [n_nb+1], t_src [nnz] = i*R + r, t_w [nnz]); ``out`` = existing buffers to refill in place."""
        lib = _lib.load()
        dev = self.device
        n = max(self.nnz, 1)
        if out is None:
            out = (torch.empty(self.n_nb + 1, dtype=torch.int32, device=dev),
                   torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev))
        t_indptr, t_src, t_w = out
        if self._t_ws is None:
            ws_bytes = lib.sg_csr_transpose_ws_bytes(self.n_seg, self.n_nb, self.nnz)
            if ws_bytes == 0:
                check(2, "sg_csr_transpose_ws_bytes")
            self._t_ws = (_bytes(ws_bytes, dev), ws_bytes, torch.empty(n, dtype=torch.int32, device=dev),
                          torch.empty(n, dtype=torch.int32, device=dev))
        ws, ws_bytes, t_perm, t_seg = self._t_ws
        check(lib.sg_csr_transpose(_p(t_indptr), _p(t_perm), _p(t_seg), _p(self.end_points), _p(self.cat_indptr),
                                   self.n_seg, self.n_nb, self.nnz, _p(ws), ws_bytes, _stream()), "sg_csr_transpose")
        check(lib.sg_multilink_transpose_finish(_p(t_src), _p(t_w), _p(t_perm), _p(t_seg), _p(self.support),
                                                self.R, self.n_dst, self.nnz, _stream()), "sg_multilink_transpose_finish")
        return out

    def transposed(self):
        """(t_indptr [n_nb+1], t_src [nnz] = i*R + r, t_w [nnz]) built once per plan."""
        if self._t is None:
            self._t = self._build_transposed()
            if self.nnz > (1 << 20) and not self.keep_transpose_scratch:
                self._t_ws = None                   # the sort scratch (~30 B/edge) is not kept for big plans
        return self._t

    def refresh_weights_(self):
        """The edge weights (``support``) were rewritten in place, the pattern did not change: re-derive only what
        depends on the weights — the transposed weight array (one launch when the sort outputs were kept,
        ``keep_transpose_scratch``; a full re-sort otherwise).  CUDA-graph capturable."""
        if self._t is not None:
            if self._t_ws is not None:
                _ws, _n, t_perm, t_seg = self._t_ws
                check(_lib.load().sg_multilink_transpose_finish(_p(self._t[1]), _p(self._t[2]), _p(t_perm), _p(t_seg),
                                                                _p(self.support), self.R, self.n_dst, self.nnz, _stream()),
                      "sg_multilink_transpose_finish")
            else:
                self._build_transposed(self._t)
        return self

    def rebuild_(self, backward=True):
        """Recompute every derived structure (schedules, transposed operands) INTO THE EXISTING BUFFERS after the
        CSR arrays were refreshed in place (load_lists_, or a device producer writing into them).  Pure kernel
        launches on the current stream with pre-allocated scratch: capturable in a CUDA graph together with the
        forward / backward that follows.  Call prepare() once before capturing."""
        if self._sched is not None:
            self._sched.rebuild_()
        if backward and self._t is not None:
            self._build_transposed(self._t)
            if self._t_sched is not None:
                self._t_sched.rebuild_()
        return self

    def t_schedule(self):
        if self.use_schedule and self._t_sched is None:
            self._t_sched = Schedule(self.transposed()[0], self.nnz, self.chunk)
        return self._t_sched

    def prepare(self, backward=True):
        """Build every derived structure now (so timed regions contain only the hot kernels)."""
        self.schedule()
        if backward:
            self.transposed()
            self.t_schedule()
        return self


# When set to a list, every fused aggregation launch appends (tag, start_event, end_event, csr) so a
# benchmark can time the gather kernels on their own stream inside its timed region.
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(tag, e0, csr):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((tag, e0, e1, csr))


class _MultiLinkAgg(torch.autograd.Function):
    """agg[i, r*D:(r+1)*D] = sum_{p in seg(r,i)} support[p] * x[end_points[p], :]  (+ wsum[i, r])."""

    @staticmethod
    def forward(ctx, x, csr):
        lib = _lib.load()
        n_nb, D = x.shape
        agg = torch.empty((csr.n_dst, csr.R * D), dtype=torch.float32, device=x.device)
        wsum = torch.empty((csr.n_dst, csr.R), dtype=torch.float32, device=x.device)
        sched = csr.schedule()
        if sched is not None:
            part = sched.partial(1, D, extra_per_row=1)
            plan, chunk, pp = _p(sched.buf), sched.chunk, _p(part)
        else:
            plan, chunk, pp = ctypes.c_void_p(0), 0, ctypes.c_void_p(0)
        e0 = _prof_begin()
        check(lib.sg_multilink_agg_fwd(_p(agg), _p(wsum), _p(x), _p(csr.support), _p(csr.end_points),
                                       _p(csr.cat_indptr), csr.R, csr.n_dst, n_nb, csr.nnz, D, plan, chunk, pp,
                                       _stream()), "sg_multilink_agg_fwd")
        _prof_end("agg_fwd", e0, csr)
        ctx.csr, ctx.D, ctx.n_nb = csr, D, n_nb
        ctx.mark_non_differentiable(wsum)
        return agg, wsum

    @staticmethod
    def backward(ctx, gagg, _gwsum):
        csr, D = ctx.csr, ctx.D
        lib = _lib.load()
        gagg = gagg.contiguous()
        gx = torch.empty((ctx.n_nb, D), dtype=torch.float32, device=gagg.device)
        t_indptr, t_src, t_w = csr.transposed()
        sched = csr.t_schedule()
        if sched is not None:
            part = sched.partial(1, D)
            plan, chunk, pp = _p(sched.buf), sched.chunk, _p(part)
        else:
            plan, chunk, pp = ctypes.c_void_p(0), 0, ctypes.c_void_p(0)
        e0 = _prof_begin()
        check(lib.sg_multilink_agg_bwd(_p(gx), _p(gagg), _p(t_w), _p(t_src), _p(t_indptr), csr.R, csr.n_dst,
                                       ctx.n_nb, csr.nnz, D, 1, plan, chunk, pp, _stream()), "sg_multilink_agg_bwd")
        _prof_end("agg_bwd", e0, csr)
        return gx, None


def _split_tf32(src, ld_dst, transpose=False):
    """(hi, lo) of a 2-D fp32 tensor, padded to ld_dst columns (sg_split_tf32)."""
    lib = _lib.load()
    rows, cols = src.shape
    if src.stride(1) != 1:
        src = src.contiguous()
    out_rows = cols if transpose else rows
    hi = torch.empty((out_rows, ld_dst), dtype=torch.float32, device=src.device)
    lo = torch.empty_like(hi)
    check(lib.sg_split_tf32(_p(hi), _p(lo), ld_dst, _p(src), rows, cols, src.stride(0), int(transpose), _stream()),
          "sg_split_tf32")
    return hi, lo


# False (shipped): the producers of the large activation operands (the gather kernel for agg, the activation-gradient
# kernel for gZ) write the TF32 hi / lo halves, the GEMM streams both.  True: the operands go to the GEMM as plain
# fp32 and are split inside the kernel (tf32x3_gemm_split_kernel) — half the operand bytes in HBM and through TMA,
# but measured SLOWER on the B200 (forward transform 0.167 vs 0.131 ms, weight gradient 0.200 vs 0.130 ms: the
# splitter's stage hand-over sits in the MMA warp's critical path and the 192 KB operand ring cannot get deeper;
# profiles/r02_summary.md).  Kept as a tested option (tests/test_gemm_gpu.py) and for A/B runs (bench.py --inkernel-split).
GEMM_INKERNEL_SPLIT = False


def _raw_ok(t):
    """Can ``t`` be a raw (plain fp32) TMA operand: 2-D, unit inner stride, 16-byte aligned rows."""
    return (GEMM_INKERNEL_SPLIT and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0
            and t.data_ptr() % 16 == 0)


def _gemm_tf32x3(D, a_hi, a_lo, b_hi, b_lo, M, N, K, mn_major=False, epilogue=0, slope=0.0, splits=1, bias=None):
    """a_lo / b_lo None: that operand is plain fp32 and is split inside the kernel."""
    lib = _lib.load()
    ws = None
    if splits > 1:
        ws = torch.empty(int(lib.sg_gemm_split_ws_bytes(M, N, splits)) // 4, dtype=torch.float32, device=D.device)
    check(lib.sg_gemm_tf32x3(_p(D), D.stride(0), _p(a_hi), _p(a_lo), a_hi.stride(0), _p(b_hi), _p(b_lo), b_hi.stride(0),
                             M, N, K, int(mn_major), int(epilogue), ctypes.c_float(slope), _p(bias), int(splits), _p(ws),
                             _stream()), "sg_gemm_tf32x3")
    return D


_SM_COUNT = {}


def _sm_count(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


FUSED_DIMS = (16, 32, 64, 128)


class _FusedAggTransform(torch.autograd.Function):
    """out = act([agg | wsum] . w_ext^T): the whole MultiLinkGCNAggregator (accum='sum') in one gather
    launch + one tcgen05 GEMM; backward = activation gradient (pre-split), two GEMMs and the transposed
    gather.  w_ext [U, R*D + R] = [W_0 | ... | W_{R-1} | b_0 ... b_{R-1}]."""

    @staticmethod
    def forward(ctx, x, w_ext, csr, slope, grad_group=None, halo=None):
        lib = _lib.load()
        ctx.grad_group = grad_group
        U, Kx = w_ext.shape
        ctx.peer = None
        if halo is not None:
            # node-partitioned run over NVLink peer memory (dist.PeerTransport): x holds this rank's OWN neighbour rows;
            # every rank stores its block into every rank's table, one flag barrier, and the gather below reads the
            # whole [n_total, D] table through the plan's global column ids
            will_backward = any(ctx.needs_input_grad[:2])
            ctx.peer = peer = halo.peer_transport(x.shape[1], U * Kx, x.device)
            ctx.n_local = x.shape[0]
            x = peer.all_gather(x, will_backward)
        n_nb, D = x.shape
        R, n_dst = csr.R, csr.n_dst
        assert Kx == R * D + R
        ld = (Kx + 31) // 32 * 32
        dev = x.device
        agg_hi = torch.empty((max(n_dst, 1), ld), dtype=torch.float32, device=dev)
        agg_lo = None if GEMM_INKERNEL_SPLIT else torch.empty_like(agg_hi)     # None: agg_hi holds plain fp32
        sched = csr.schedule()
        if sched is not None:
            part = sched.partial(1, D, extra_per_row=1)
            plan, chunk, pp = _p(sched.buf), sched.chunk, _p(part)
        else:
            plan, chunk, pp = ctypes.c_void_p(0), 0, ctypes.c_void_p(0)
        e0 = _prof_begin()
        check(lib.sg_multilink_agg_fwd_split(_p(agg_hi), _p(agg_lo), ld, _p(x), _p(csr.support), _p(csr.end_points),
                                             _p(csr.cat_indptr), R, n_dst, n_nb, csr.nnz, D, plan, chunk, pp,
                                             _stream()), "sg_multilink_agg_fwd_split")
        _prof_end("agg_fwd", e0, csr)
        w_hi, w_lo = _split_tf32(w_ext, ld)
        out = torch.empty((n_dst, U), dtype=torch.float32, device=dev)
        if n_dst > 0:
            e0 = _prof_begin()
            _gemm_tf32x3(out, agg_hi, agg_lo, w_hi, w_lo, n_dst, U, Kx, epilogue=1, slope=slope)
            _prof_end("gemm_fwd", e0, csr)
        ctx.csr, ctx.slope, ctx.dims = csr, slope, (n_nb, D, R, n_dst, U, Kx, ld)
        ctx.raw = agg_lo is None
        ctx.save_for_backward(agg_hi, agg_lo, out, w_ext)
        if ctx.peer is not None and not will_backward:
            ctx.peer.release()          # forward-only: no backward barrier will cover the table's readers
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        agg_hi, agg_lo, out, w_ext = ctx.saved_tensors
        csr, slope = ctx.csr, ctx.slope
        n_nb, D, R, n_dst, U, Kx, ld = ctx.dims
        dev = gout.device
        gout = gout.contiguous()
        ldz = (U + 3) // 4 * 4
        gz_hi = torch.empty((max(n_dst, 1), ldz), dtype=torch.float32, device=dev)
        gz_lo = None if ctx.raw else torch.empty_like(gz_hi)
        check(lib.sg_act_bwd_split(_p(gz_hi), _p(gz_lo), ldz, _p(gout), _p(out), n_dst, U, ctypes.c_float(slope),
                                   _stream()), "sg_act_bwd_split")
        gx = gw = None
        peer = ctx.peer
        if ctx.needs_input_grad[1]:
            n_gw = (U * Kx + 3) // 4 * 4      # flat, padded to 16 bytes: the peer transport moves float4
            gw_flat = torch.zeros(n_gw, dtype=torch.float32, device=dev) if (n_dst == 0 or n_gw != U * Kx) else \
                torch.empty(n_gw, dtype=torch.float32, device=dev)
            gw = gw_flat[:U * Kx].view(U, Kx)
            if n_dst > 0:
                tiles = ((U + 127) // 128) * ((Kx + 255) // 256)
                kb = (n_dst + 31) // 32
                splits = max(1, min(_sm_count(dev) // tiles, kb // 4))
                e0 = _prof_begin()
                _gemm_tf32x3(gw, gz_hi, gz_lo, agg_hi, agg_lo, U, Kx, n_dst, mn_major=True, splits=splits)
                _prof_end("gemm_dw", e0, csr)
            if peer is not None and ctx.grad_group is not None:
                peer.push_grad(gw_flat)   # into every rank's slot; summed after the backward barrier below
            elif ctx.grad_group is not None:
                # partitioned run: sum the packed weight gradient over ranks here, ONE collective per layer
                # direction, issued before the transposed gather so NCCL overlaps it
                import torch.distributed as dist
                if dist.get_backend(ctx.grad_group) == "nccl":
                    dist.all_reduce(gw, group=ctx.grad_group)
                else:
                    host = gw.cpu()
                    dist.all_reduce(host, group=ctx.grad_group)
                    gw.copy_(host)
        if ctx.needs_input_grad[0]:
            if peer is None:
                gx = torch.empty((n_nb, D), dtype=torch.float32, device=dev)
            if n_dst > 0:
                wt_hi, wt_lo = _split_tf32(w_ext[:, :R * D], ldz, transpose=True)       # [R*D, ldz]
                gagg = torch.empty((n_dst, R * D), dtype=torch.float32, device=dev)
                e0 = _prof_begin()
                _gemm_tf32x3(gagg, gz_hi, gz_lo, wt_hi, wt_lo, n_dst, R * D, U)
                _prof_end("gemm_dagg", e0, csr)
            else:
                gagg = torch.zeros((0, R * D), dtype=torch.float32, device=dev)
            t_indptr, t_src, t_w = csr.transposed()
            sched = csr.t_schedule()
            if sched is not None:
                part = sched.partial(1, D)
                plan, chunk, pp = _p(sched.buf), sched.chunk, _p(part)
            else:
                plan, chunk, pp = ctypes.c_void_p(0), 0, ctypes.c_void_p(0)
            e0 = _prof_begin()
            if peer is not None:
                # reduce-scatter fused into its producer: every finished gradient row goes straight into the owner's
                # staging slot of this rank (NVLink stores inside the gather launch)
                stage, owner_lo, world = peer.scatter_args()
                check(lib.sg_multilink_agg_bwd_peer(stage, owner_lo, world, _p(gagg), _p(t_w), _p(t_src), _p(t_indptr), R,
                                                    n_dst, n_nb, csr.nnz, D, plan, chunk, pp, _stream()),
                      "sg_multilink_agg_bwd_peer")
            else:
                check(lib.sg_multilink_agg_bwd(_p(gx), _p(gagg), _p(t_w), _p(t_src), _p(t_indptr), R, n_dst, n_nb, csr.nnz,
                                               D, 1, plan, chunk, pp, _stream()), "sg_multilink_agg_bwd")
            _prof_end("agg_bwd", e0, csr)
        if peer is not None:
            peer.barrier()                    # every rank's rows and gradient slots have landed
            peer._open = False
            if ctx.needs_input_grad[0]:
                gx = peer.reduce_rows()       # [n_local, D]: the world slots summed in rank order
            if gw is not None and ctx.grad_group is not None:
                peer.reduce_grad(gw_flat)
        return gx, gw, None, None, None, None


class _PackWExt(torch.autograd.Function):
    """w_ext [U, R*D + R] = [W_0 | ... | W_{R-1} | b_0 ... b_{R-1}] from the per-level parameters.

    Same values as ``torch.cat(ws + [torch.stack(bs, 1)], 1)``; what differs is the backward: autograd's
    cat/stack backward hands every parameter a strided slice that AccumulateGrad then clones — 2R+1 tiny
    launches per layer direction at the tail of the step.  Here the packed gradient is regrouped by TWO
    launches into a level-major buffer whose contiguous rows ``gW[r]`` / ``gb[r]`` become the parameter
    gradients directly."""

    @staticmethod
    def forward(ctx, R, *params):
        ws, bs = params[:R], params[R:]
        ctx.dims = (R,) + tuple(ws[0].shape)
        return torch.cat(list(ws) + [torch.stack(bs, dim=1)], dim=1)

    @staticmethod
    def backward(ctx, g):
        R, U, D = ctx.dims
        gw = g[:, :R * D].reshape(U, R, D).permute(1, 0, 2).contiguous()    # [R, U, D]
        gb = g[:, R * D:].t().contiguous()                                  # [R, U]
        return (None,) + tuple(gw[r] for r in range(R)) + tuple(gb[r] for r in range(R))


def pack_w_ext(ws, bs):
    """[W_0 | ... | W_{R-1} | b_0 ... b_{R-1}] — the B operand of the fused transform (see _PackWExt)."""
    return _PackWExt.apply(len(ws), *ws, *bs)


def fused_agg_transform(x, w_ext, csr, slope, grad_group=None, halo=None):
    """act([agg | wsum] . w_ext^T) with act = leaky(slope) (slope 1.0 = identity, 0.0 = ReLU).

    ``halo``: a ``dist.HaloPlan`` of mode 'peer' — ``x`` then holds only this rank's OWN neighbour rows and the
    exchange (all-gather forward, reduce-scatter of the data gradient and all-reduce of the weight gradient
    backward) runs inside this op over NVLink peer memory."""
    if x.dtype != torch.float32 or not x.is_cuda or x.dim() != 2:
        raise TypeError("x must be a 2-D float32 CUDA tensor")
    if halo is not None:
        if halo.mode != "peer":
            raise ValueError("halo must be a HaloPlan of mode 'peer' (other modes go through dist.halo_exchange)")
        if x.shape[0] != halo.n_local or csr.n_nb != halo.n_ext:
            raise ValueError(f"x must hold the rank's {halo.n_local} own rows and the plan index all {halo.n_ext} rows")
    elif x.shape[0] != csr.n_nb:
        raise ValueError(f"x has {x.shape[0]} rows but the plan indexes {csr.n_nb} neighbour rows")
    if x.shape[1] not in FUSED_DIMS:
        raise ValueError(f"the fused path needs D in {FUSED_DIMS}")
    return _FusedAggTransform.apply(x.contiguous(), w_ext, csr, float(slope), grad_group, halo)


def multilink_aggregate(x, csr):
    """Fused all-relations neighbour aggregation; returns (agg [n_dst, R*D], wsum [n_dst, R])."""
    if x.dtype != torch.float32 or not x.is_cuda or x.dim() != 2:
        raise TypeError("x must be a 2-D float32 CUDA tensor")
    if x.shape[0] != csr.n_nb:
        raise ValueError(f"x has {x.shape[0]} rows but the plan indexes {csr.n_nb} neighbour rows")
    return _MultiLinkAgg.apply(x.contiguous(), csr)
