"""Multi-tensor clip_global_norm + Adam against a numpy restatement of MXNet's documented
update rule (gluon.utils.clip_global_norm, mx.optimizer.Adam / adam_update).  MXNet is not vendored in the
reference tree: parity unpinned beyond those documented semantics."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def np_clip(grads, max_norm):
    norm = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    scale = max_norm / (norm + 1e-8)
    return norm, ([g * np.float32(scale) for g in grads] if scale < 1.0 else grads)


def np_adam(w, g, m, v, t, lr, wd, b1=0.9, b2=0.999, eps=1e-8):
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    g = g + wd * w
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    return w - lr_t * m / (np.sqrt(v) + eps), m, v


@pytest.mark.parametrize("max_norm", [0.5, 1e6])
def test_clip_and_adam_match_mxnet_rule(max_norm):
    from stargcn_b200.optim import FusedAdam
    rs = np.random.RandomState(0)
    shapes = [(250, 64)] * 10 + [(250,)] * 10 + [(75, 250), (75,), (64, 75), (64,), (70000, 64), (1,)]
    ws = [rs.normal(size=s).astype(np.float32) for s in shapes]
    params = [torch.nn.Parameter(torch.from_numpy(w.copy()).cuda()) for w in ws]
    opt = FusedAdam(params, learning_rate=1e-2, wd=1e-4)
    m = [np.zeros_like(w, dtype=np.float64) for w in ws]
    v = [np.zeros_like(w, dtype=np.float64) for w in ws]
    w64 = [w.astype(np.float64) for w in ws]
    for t in range(1, 4):
        gs = [rs.normal(size=s).astype(np.float32) * 0.1 for s in shapes]
        for p, g in zip(params, gs):
            p.grad = torch.from_numpy(g.copy()).cuda()
        norm = opt.clip_global_norm(max_norm)
        opt.step(1.0)
        ref_norm, gs_c = np_clip(gs, max_norm)
        assert abs(float(norm) - ref_norm) <= 1e-5 * ref_norm
        for k in range(len(ws)):
            w64[k], m[k], v[k] = np_adam(w64[k], gs_c[k].astype(np.float64), m[k], v[k], t, 1e-2, 1e-4)
            assert rel_err(params[k].detach().cpu().numpy(), w64[k]) <= 1e-5
            assert rel_err(params[k].grad.cpu().numpy(), gs_c[k]) <= 1e-6      # gradients rescaled in place
    # bit-identical reruns of the norm (fixed-order reduction)
    n1 = float(opt.clip_global_norm(1.0)); n2 = float(opt.clip_global_norm(1.0))
    assert n1 == n2
