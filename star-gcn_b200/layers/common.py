"""Activations and the deferred-shape Dense layer used by the layer mirror.

``get_activation`` follows mxgraph/layers/common.py:32-57 ('leaky' = LeakyReLU(0.1), 'elu',
'identity', relu/sigmoid/tanh/softrelu/softsign, None -> identity, modules pass through).
``Dense`` stands in for ``gluon.nn.Dense(units, flatten=False)`` with MXNet's deferred input
size and the initialiser the experiment script uses, Xavier(factor_type='in') uniform
(experiments/STAR-GCN.py:548): U(-sqrt(3/fan_in), sqrt(3/fan_in)), zero bias.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.lazy import LazyModuleMixin


class IdentityActivation(nn.Module):
    def forward(self, x):
        return x


class ELU(nn.Module):
    """-alpha * relu(1 - exp(x)) + relu(x)   (common.py:9-29)"""

    def __init__(self, alpha=1.0):
        super().__init__()
        self._alpha = alpha

    def forward(self, x):
        return -self._alpha * F.relu(1.0 - torch.exp(x)) + F.relu(x)


class _Softsign(nn.Module):
    def forward(self, x):
        return F.softsign(x)


def get_activation(act):
    if act is None:
        return IdentityActivation()
    if isinstance(act, str):
        if act == "leaky":
            return nn.LeakyReLU(0.1)
        if act == "identity":
            return IdentityActivation()
        if act == "elu":
            return ELU()
        table = {"relu": nn.ReLU, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "softrelu": nn.Softplus,
                 "softsign": _Softsign}
        if act in table:
            return table[act]()
        raise NotImplementedError(act)
    return act


def activation_code(act):
    """SG_ACT_* code if the activation can ride in a fused epilogue, else None."""
    if act is None or isinstance(act, IdentityActivation):
        return 0
    if isinstance(act, nn.LeakyReLU) and abs(act.negative_slope - 0.1) < 1e-12:
        return 1
    if isinstance(act, nn.ReLU):
        return 2
    return None


def xavier_in_uniform_(weight, fan_in):
    bound = math.sqrt(3.0 / max(fan_in, 1))
    with torch.no_grad():
        weight.uniform_(-bound, bound)
    return weight


class Dense(LazyModuleMixin, nn.Module):
    """y = x W^T + b with W (units, in_units); in_units inferred at first call (MXNet's deferred init).

    Until then ``weight`` / ``bias`` are registered ``UninitializedParameter`` s (torch's lazy-module protocol):
    ``state_dict()`` works on a freshly built model and ``load_state_dict()`` materialises them from the
    checkpoint's shapes — the counterpart of gluon's ``load_parameters`` on a deferred-init block."""

    cls_to_become = None

    def __init__(self, units, in_units=None, use_bias=True, device=None):
        super().__init__()
        self._units = units
        self._use_bias = use_bias
        self.weight = nn.UninitializedParameter()
        if use_bias:
            self.bias = nn.UninitializedParameter()
        else:
            self.register_parameter("bias", None)
        if in_units is not None:
            self._materialize(in_units, device)

    def _materialize(self, in_units, device):
        if isinstance(self.weight, nn.UninitializedParameter):
            self.weight.materialize((self._units, in_units), device=device, dtype=torch.float32)
            xavier_in_uniform_(self.weight, in_units)
        if self._use_bias and isinstance(self.bias, nn.UninitializedParameter):
            self.bias.materialize((self._units,), device=device, dtype=torch.float32)
            with torch.no_grad():
                self.bias.zero_()

    def initialize_parameters(self, x, *args, **kwargs):
        if self.has_uninitialized_params():
            self._materialize(x.shape[-1], x.device)

    def forward(self, x, act=None):
        """x W^T + b on the tcgen05 GEMM; ``act`` in {None, 'leaky', 'relu'} rides in its epilogue."""
        from ..decoder import fused_dense
        if self.has_uninitialized_params():       # direct call that bypassed the lazy pre-hook
            self._materialize(x.shape[-1], x.device)
        lead = x.shape[:-1]
        y = fused_dense(x.reshape(-1, x.shape[-1]), self.weight, self.bias, act)
        return y.reshape(*lead, self._units)
