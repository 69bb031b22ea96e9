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


def _as_np(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dtype)


class MultiLinkCSR:
    """R rating-level CSR neighbour lists over the same n_dst destination nodes.

    Parameters mirror the aggregator call (aggregators.py:111-128): ``end_points_l[r]`` are
    LOCAL row ids into the neighbour feature matrix, ``indptr_l[r]`` has n_dst+1 entries,
    ``support_l[r]`` is the edge normalisation 1/sqrt(d_i d_j).  Levels without edges may
    arrive as the reference's length-1 dummies (``empty_as_zero``, graph.py:221-222): only
    the first ``indptr_l[r][-1]`` entries of each list are used.
    """

    def __init__(self, end_points_l, indptr_l, support_l, n_nb, device=None, chunk=DEFAULT_CHUNK,
                 use_schedule=True):
        if not (len(end_points_l) == len(indptr_l) == len(support_l)) or len(indptr_l) == 0:
            raise ValueError("end_points_l, indptr_l and support_l must be non-empty lists of equal length")
        if device is None:
            device = next((t.device for t in list(end_points_l) + list(indptr_l)
                           if isinstance(t, torch.Tensor) and t.is_cuda), torch.device("cuda"))
        self.device = torch.device(device)
        self.R = len(indptr_l)
        ptrs = [_as_np(p, np.int32) for p in indptr_l]
        self.n_dst = int(ptrs[0].shape[0]) - 1
        if any(p.shape[0] != self.n_dst + 1 for p in ptrs):
            raise ValueError("every indptr must have n_dst + 1 entries")
        self.nnz_l = [int(p[-1]) for p in ptrs]
        self.nnz = int(sum(self.nnz_l))
        self.n_nb = int(n_nb)
        eps = [_as_np(e, np.int32)[:n] for e, n in zip(end_points_l, self.nnz_l)]
        sup = [_as_np(s, np.float32)[:n] for s, n in zip(support_l, self.nnz_l)]
        offs = np.concatenate([[0], np.cumsum(self.nnz_l)]).astype(np.int64)
        cat_ptr = np.concatenate([ptrs[0][:1].astype(np.int64)] +
                                 [ptrs[r][1:].astype(np.int64) + offs[r] for r in range(self.R)])
        if cat_ptr[-1] != self.nnz or self.R * self.n_dst >= 2 ** 31:
            raise ValueError("inconsistent indptr lists")
        ep_cat = np.concatenate(eps) if self.nnz else np.zeros(0, np.int32)
        if self.nnz and (ep_cat.min() < 0 or ep_cat.max() >= self.n_nb):
            raise ValueError("end point index out of range of the neighbour feature matrix")
        self._init_device(torch.from_numpy(ep_cat), torch.from_numpy(np.concatenate(sup) if self.nnz else
                                                                     np.zeros(0, np.float32)),
                          torch.from_numpy(cat_ptr.astype(np.int32)), chunk, use_schedule)

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
        self._t_sched = None
        self.h2d_bytes = 4 * (2 * self.nnz + self.n_seg + 1)

    def schedule(self):
        if self.use_schedule and self._sched is None:
            self._sched = Schedule(self.cat_indptr, self.nnz, self.chunk)
        return self._sched

    def transposed(self):
        """(t_indptr [n_nb+1], t_src [nnz] = i*R + r, t_w [nnz]) built once per plan."""
        if self._t is None:
            lib = _lib.load()
            dev = self.device
            t_indptr = torch.empty(self.n_nb + 1, dtype=torch.int32, device=dev)
            n = max(self.nnz, 1)
            t_perm = torch.empty(n, dtype=torch.int32, device=dev)
            t_seg = torch.empty(n, dtype=torch.int32, device=dev)
            ws_bytes = lib.sg_csr_transpose_ws_bytes(self.n_seg, self.n_nb, self.nnz)
            if ws_bytes == 0:
                check(2, "sg_csr_transpose_ws_bytes")
            ws = _bytes(ws_bytes, dev)
            check(lib.sg_csr_transpose(_p(t_indptr), _p(t_perm), _p(t_seg), _p(self.end_points), _p(self.cat_indptr),
                                       self.n_seg, self.n_nb, self.nnz, _p(ws), ws_bytes, _stream()),
                  "sg_csr_transpose")
            t_src = torch.empty(n, dtype=torch.int32, device=dev)
            t_w = torch.empty(n, dtype=torch.float32, device=dev)
            check(lib.sg_multilink_transpose_finish(_p(t_src), _p(t_w), _p(t_perm), _p(t_seg), _p(self.support),
                                                    self.R, self.n_dst, self.nnz, _stream()),
                  "sg_multilink_transpose_finish")
            self._t = (t_indptr, t_src, t_w)
        return self._t

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


def multilink_aggregate(x, csr):
    """Fused all-relations neighbour aggregation; returns (agg [n_dst, R*D], wsum [n_dst, R])."""
    if x.dtype != torch.float32 or not x.is_cuda or x.dim() != 2:
        raise TypeError("x must be a 2-D float32 CUDA tensor")
    if x.shape[0] != csr.n_nb:
        raise ValueError(f"x has {x.shape[0]} rows but the plan indexes {csr.n_nb} neighbour rows")
    return _MultiLinkAgg.apply(x.contiguous(), csr)
