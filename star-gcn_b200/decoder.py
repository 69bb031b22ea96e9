"""Masked-embedding reconstruction decoder and rating head of STAR-GCN on the sm_100a kernels.

Mirrors, on torch CUDA tensors (reference: /root/reference/experiments/STAR-GCN.py):

  get_embed(table, node_ids, embed_noise, use_mask)   Net.get_embed :264-300
  take_rows(x, idx)                                   mx.nd.take(x, idx)  :429-432, :447, :454
  fused_dense(x, W, b, act)                           gluon nn.Dense(flatten=False) (+ activation)
  EmbedMap(units, act)                                embed_maps[block][key] = Dense -> act -> Dense :237-245
  recon_loss(gt, pred)                                mean(sum(square(gt - pred), -1))  :625
  inner_product(a, b)                                 InnerProductLayer (mxgraph/layers/layers.py:217-222)
  l2_loss(pred, label)                                gluon.loss.L2Loss(...).mean()  :611-616

Dense layers run on the tcgen05 3xTF32 GEMM (csrc/gemm.cu) with bias + LeakyReLU in its epilogue;
row gathers/scatters reuse the CSR gather kernels (deterministic, no atomics); reductions are
fixed-order.  No CPU fallback.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib, seg_op
from ._lib import check
from . import graph as _graph
from .graph import _gemm_tf32x3, _sm_count, _split_tf32
from .seg_op import _p, _stream


def _chk2d(x, name):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError(f"{name} must be a 2-D float32 tensor, got {x.dtype} {tuple(x.shape)}")
    return x.contiguous()


# ------------------------------------------------------------------------------------------------
# Dense on the tensor cores
# ------------------------------------------------------------------------------------------------
class _FusedDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, slope):
        n, K = x.shape
        N = weight.shape[0]
        dev = x.device
        out = torch.empty((n, N), dtype=torch.float32, device=dev)
        ldk = (K + 3) // 4 * 4
        x_hi = x_lo = None
        if n > 0:
            if _graph._raw_ok(x):              # plain fp32 operand, split inside the kernel
                x_hi = x
            else:
                x_hi, x_lo = _split_tf32(x, ldk)
            w_hi, w_lo = _split_tf32(weight, ldk)
            _gemm_tf32x3(out, x_hi, x_lo, w_hi, w_lo, n, N, K, epilogue=0 if slope == 1.0 else 1, slope=slope,
                         bias=bias)
        ctx.slope, ctx.dims, ctx.has_bias = slope, (n, K, N, ldk), bias is not None
        ctx.save_for_backward(x_hi, x_lo, weight, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        x_hi, x_lo, weight, out = ctx.saved_tensors
        n, K, N, ldk = ctx.dims
        dev = gout.device
        gx = gw = gb = None
        if n == 0:
            if ctx.needs_input_grad[0]:
                gx = torch.zeros((0, K), dtype=torch.float32, device=dev)
            if ctx.needs_input_grad[1]:
                gw = torch.zeros_like(weight)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                gb = torch.zeros(N, dtype=torch.float32, device=dev)
            return gx, gw, gb, None
        gout = gout.contiguous()
        ldz = (N + 3) // 4 * 4
        gz_hi = torch.empty((n, ldz), dtype=torch.float32, device=dev)
        # x_lo None = the forward took x raw; gZ then goes raw as well (a raw B operand needs a raw A)
        gz_lo = None if x_lo is None else torch.empty_like(gz_hi)
        check(lib.sg_act_bwd_split(_p(gz_hi), _p(gz_lo), ldz, _p(gout), _p(out), n, N, ctypes.c_float(ctx.slope),
                                   _stream()), "sg_act_bwd_split")
        if ctx.needs_input_grad[0]:
            wt_hi, wt_lo = _split_tf32(weight, ldz, transpose=True)            # [K, ldz]
            gx = torch.empty((n, K), dtype=torch.float32, device=dev)
            _gemm_tf32x3(gx, gz_hi, gz_lo, wt_hi, wt_lo, n, K, N)
        if ctx.needs_input_grad[1]:
            gw = torch.empty((N, K), dtype=torch.float32, device=dev)
            tiles = ((N + 127) // 128) * ((K + 255) // 256)
            splits = max(1, min(_sm_count(dev) // tiles, ((n + 31) // 32) // 4))
            _gemm_tf32x3(gw, gz_hi, gz_lo, x_hi, x_lo, N, K, n, mn_major=True, splits=splits)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = torch.empty(N, dtype=torch.float32, device=dev)
            ws = torch.empty(int(lib.sg_colsum_ws_bytes(N)) // 4, dtype=torch.float32, device=dev)
            check(lib.sg_colsum(_p(gb), _p(gz_hi), _p(gz_lo), n, N, ldz, _p(ws), _stream()), "sg_colsum")
        return gx, gw, gb, None


_SLOPES = {None: 1.0, "identity": 1.0, "leaky": 0.1, "relu": 0.0}


def fused_dense(x, weight, bias=None, act=None):
    """act(x W^T + b), W (units, in_units); act in {None, 'identity', 'leaky', 'relu'}."""
    if act not in _SLOPES:
        raise NotImplementedError(f"activation {act!r} cannot ride in the GEMM epilogue")
    x = _chk2d(x, "x")
    if weight.shape[1] != x.shape[1]:
        raise ValueError(f"weight {tuple(weight.shape)} does not match input width {x.shape[1]}")
    return _FusedDense.apply(x, weight.contiguous(), bias, _SLOPES[act])


# ------------------------------------------------------------------------------------------------
# D1  masked embedding lookup
# ------------------------------------------------------------------------------------------------
def _arange_indptr(n, device):
    return torch.arange(n + 1, dtype=torch.int32, device=device)


class _MaskedEmbed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, node_ids, noise):
        n, (n_table, D) = node_ids.numel(), table.shape
        out = torch.empty((n, D), dtype=torch.float32, device=table.device)
        eff = torch.empty(max(n, 1), dtype=torch.int32, device=table.device)[:n]
        check(_lib.load().sg_masked_embed_fwd(_p(out), _p(eff), _p(table), _p(node_ids), _p(noise), n, n_table, D,
                                              _stream()), "sg_masked_embed_fwd")
        ctx.eff, ctx.shape = eff, (n_table, D)
        return out

    @staticmethod
    def backward(ctx, gout):
        # gtable[id', :] += gout[i, :] for every unmasked i: a gather over the stable transpose of the
        # (one edge per row) pattern — duplicates are summed in ascending i, no atomics
        n_table, D = ctx.shape
        eff = ctx.eff
        n = eff.numel()
        if n == 0:
            return torch.zeros((n_table, D), dtype=torch.float32, device=gout.device), None, None
        mask = (eff >= 0)
        pat = seg_op.CSRPattern(eff.clamp_min(0), _arange_indptr(n, eff.device), n_table)
        g = seg_op._weighted_pool_bwd_data(gout.contiguous().unsqueeze(0), mask.to(torch.float32).unsqueeze(0), pat,
                                           n_table)
        return g[0], None, None


def get_embed(table, node_ids, embed_noise=None, use_mask=True):
    """Net.get_embed for one node type: ``table`` (N_all, D) embedding weight, ``node_ids`` int32 ids,
    ``embed_noise`` (N_all,) int32 with -1 = mask to zero, i = use row i."""
    table = _chk2d(table, "table")
    node_ids = seg_op._chk_int(node_ids, "node_ids")
    noise = None
    if use_mask:
        if embed_noise is None:
            raise ValueError("use_mask=True needs embed_noise")
        noise = seg_op._chk_int(embed_noise, "embed_noise")
        if noise.numel() != table.shape[0]:
            raise ValueError("embed_noise must have one entry per table row")
    return _MaskedEmbed.apply(table, node_ids, noise)


def take_rows(x, idx):
    """mx.nd.take(x, idx) on axis 0 (differentiable; duplicate indices accumulate deterministically)."""
    x = _chk2d(x, "x")
    idx = seg_op._chk_int(idx, "idx")
    n = idx.numel()
    out = seg_op.seg_pool(x.unsqueeze(0), idx, _arange_indptr(n, x.device), pool_type="sum")
    return out[0]


class EmbedMap(nn.Module):
    """embed_maps[block][key]: Dense(units) -> activation -> Dense(units) applied to take(h, idx)."""

    def __init__(self, units, act="leaky", in_units=None):
        super().__init__()
        from .layers.common import Dense
        if act not in _SLOPES:
            raise NotImplementedError(act)
        self._act = act
        self.l0 = Dense(units, in_units=in_units)
        self.l1 = Dense(units, in_units=units if in_units is not None else None)

    def forward(self, h, idx=None):
        if idx is not None:
            h = take_rows(h, idx)
        return self.l1(self.l0(h, act=self._act))


# ------------------------------------------------------------------------------------------------
# D3 / D4  losses and the inner-product rating head
# ------------------------------------------------------------------------------------------------
class _SqErr(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, scale):
        lib = _lib.load()
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        ws = torch.empty(int(lib.sg_reduce_ws_bytes()) // 4, dtype=torch.float32, device=a.device)
        check(lib.sg_sq_err_fwd(_p(loss), _p(a), _p(b), a.numel(), ctypes.c_float(scale), _p(ws), _stream()),
              "sg_sq_err_fwd")
        ctx.scale = scale
        ctx.save_for_backward(a, b)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        a, b = ctx.saved_tensors
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        check(_lib.load().sg_sq_err_bwd(_p(ga), _p(gb), _p(a), _p(b), _p(gloss.contiguous()), a.numel(),
                                        ctypes.c_float(ctx.scale), _stream()), "sg_sq_err_bwd")
        return ga, gb, None


def recon_loss(gt_emb, pred_emb):
    """mean over nodes of the summed squared error (STAR-GCN.py:625).  Gradient flows into BOTH
    arguments: the reference does not detach gt_embeddings (:359-363)."""
    gt_emb, pred_emb = _chk2d(gt_emb, "gt_emb"), _chk2d(pred_emb, "pred_emb")
    if gt_emb.shape != pred_emb.shape:
        raise ValueError("gt_emb and pred_emb must have the same shape")
    return _SqErr.apply(gt_emb, pred_emb, 1.0 / max(gt_emb.shape[0], 1))


def l2_loss(pred, label):
    """gluon.loss.L2Loss()(pred, label).mean() = mean(0.5 (pred - label)^2)."""
    pred, label = pred.reshape(-1, 1), label.reshape(-1, 1)
    pred, label = _chk2d(pred, "pred"), _chk2d(label, "label")
    return _SqErr.apply(pred, label, 0.5 / max(pred.shape[0], 1))


class _RowDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        n, D = a.shape
        out = torch.empty((n, 1), dtype=torch.float32, device=a.device)
        check(_lib.load().sg_rowdot_fwd(_p(out), _p(a), _p(b), n, D, _stream()), "sg_rowdot_fwd")
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        n, D = a.shape
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        check(_lib.load().sg_rowdot_bwd(_p(ga), _p(gb), _p(gout.contiguous()), _p(a), _p(b), n, D, _stream()),
              "sg_rowdot_bwd")
        return ga, gb


def inner_product(a, b):
    """sum(a * b, axis=1, keepdims=True) — InnerProductLayer's reduction."""
    a, b = _chk2d(a, "data1"), _chk2d(b, "data2")
    if a.shape != b.shape:
        raise ValueError("inner_product needs equal shapes")
    return _RowDot.apply(a, b)


__all__ = ["fused_dense", "get_embed", "take_rows", "EmbedMap", "recon_loss", "l2_loss", "inner_product"]
