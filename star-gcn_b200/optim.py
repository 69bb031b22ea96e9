"""Global-norm gradient clipping + Adam over all parameters in two launches.

Mirrors what the reference's training loop does per iteration (experiments/STAR-GCN.py:630-632):
``params_clip_global_norm(net.collect_params(), GRAD_CLIP, ctx)`` (mxgraph/utils.py:104-107, i.e.
``gluon.utils.clip_global_norm``) followed by ``trainer.step(1.0)`` with ``gluon.Trainer(..., 'adam',
{'learning_rate': LR, 'wd': WD})`` (:552-553).  MXNet's Adam: ``lr_t = lr sqrt(1-b2^t)/(1-b1^t)``,
``g += wd w``, ``m = b1 m + (1-b1) g``, ``v = b2 v + (1-b2) g^2``, ``w -= lr_t m / (sqrt(v) + eps)``.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import check
from .seg_op import _p, _stream


class FusedAdam:
    def __init__(self, params, learning_rate=1e-3, wd=0.0, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam needs at least one parameter")
        if any(isinstance(p, torch.nn.UninitializedParameter) for p in self.params):
            raise ValueError("FusedAdam: a parameter is still uninitialised (deferred shape) — run one forward pass "
                             "or load a checkpoint before building the optimiser")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("FusedAdam handles contiguous float32 CUDA parameters only")
        self.lr, self.wd, self.beta1, self.beta2, self.eps = learning_rate, wd, beta1, beta2, epsilon
        self.t = 0
        dev = self.params[0].device
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        chunk = int(_lib.load().sg_optim_chunk())
        work = [(t, c) for t, p in enumerate(self.params) for c in range((p.numel() + chunk - 1) // chunk)]
        self.n_work = len(work)
        self._work = torch.tensor(work, dtype=torch.int32, device=dev).contiguous()
        self._numels = torch.tensor([p.numel() for p in self.params], dtype=torch.int64, device=dev)
        self._ptr = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)
        self._p_ptrs, self._m_ptrs, self._v_ptrs = self._ptr(self.params), self._ptr(self.m), self._ptr(self.v)
        self._p_key = tuple(p.data_ptr() for p in self.params)
        self._g_ptrs, self._g_key = None, None
        self._ws = torch.empty(max(self.n_work, 1), dtype=torch.float32, device=dev)
        self._norm = torch.zeros(2, dtype=torch.float32, device=dev)
        self._clipped = False

    @property
    def learning_rate(self):
        return self.lr

    def set_learning_rate(self, lr):
        self.lr = lr

    def _grad_table(self):
        # parameters that were moved / re-materialised since the last step (model.to(...), load_state_dict into new
        # storage): refresh the pointer table instead of updating stale memory
        p_key = tuple(p.data_ptr() for p in self.params)
        if p_key != self._p_key:
            for p, m in zip(self.params, self.m):
                if p.shape != m.shape or p.device != m.device:
                    raise RuntimeError("FusedAdam: a parameter changed shape or device after the optimiser was built")
            self._p_ptrs, self._p_key = self._ptr(self.params), p_key
        grads = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("every parameter needs a gradient before clip / step")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            grads.append(p.grad)
        key = tuple(g.data_ptr() for g in grads)
        if key != self._g_key:                      # .grad buffers were reallocated (p.grad = None between steps)
            self._g_ptrs, self._g_key = self._ptr(grads), key
        return self._g_ptrs

    def clip_global_norm(self, max_norm):
        """Returns the global gradient norm as a device scalar; the rescaling itself is folded into ``step``
        (the gradient arrays are also rewritten there, as clip_global_norm does in place)."""
        g = self._grad_table()
        check(_lib.load().sg_global_norm(_p(self._norm), _p(g), _p(self._numels), _p(self._work), self.n_work,
                                         ctypes.c_float(max_norm), _p(self._ws), _stream()), "sg_global_norm")
        self._clipped = True
        return self._norm[0]

    def step(self, batch_size=1.0):
        g = self._grad_table()
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        check(_lib.load().sg_multi_adam(_p(self._p_ptrs), _p(g), _p(self._m_ptrs), _p(self._v_ptrs), _p(self._numels),
                                        _p(self._work), self.n_work, ctypes.c_float(lr_t), ctypes.c_float(self.beta1),
                                        ctypes.c_float(self.beta2), ctypes.c_float(self.eps), ctypes.c_float(self.wd),
                                        ctypes.c_float(1.0 / batch_size), _p(self._norm) if self._clipped else None,
                                        1, _stream()), "sg_multi_adam")
        self._clipped = False


__all__ = ["FusedAdam"]
