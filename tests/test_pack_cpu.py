"""pack_w_ext: the packed B operand of the fused transform and its two-launch gradient regrouping
must equal autograd's own cat/stack (values, gradients, and gradient accumulation over two uses)."""
import pytest
import torch

import stargcn_b200  # noqa: F401
from stargcn_b200.graph import pack_w_ext


@pytest.mark.parametrize("R,U,D", [(1, 4, 16), (5, 250, 64), (10, 250, 64), (3, 7, 5)])
def test_pack_w_ext_matches_cat_stack(R, U, D):
    g = torch.Generator().manual_seed(R * 1000 + U)
    ws = [torch.randn(U, D, generator=g, requires_grad=True) for _ in range(R)]
    bs = [torch.randn(U, generator=g, requires_grad=True) for _ in range(R)]
    ws2 = [w.detach().clone().requires_grad_(True) for w in ws]
    bs2 = [b.detach().clone().requires_grad_(True) for b in bs]
    a = pack_w_ext(ws, bs)
    b = torch.cat(ws2 + [torch.stack(bs2, dim=1)], dim=1)
    assert a.shape == (U, R * D + R) and torch.equal(a, b)
    gout = torch.randn(a.shape, generator=g)
    a.backward(gout)
    b.backward(gout)
    for p, q in zip(ws + bs, ws2 + bs2):
        assert p.grad.is_contiguous() and torch.equal(p.grad, q.grad)


def test_pack_w_ext_ordinal_sharing_chain_and_accumulation():
    """Cumulative (ordinal-sharing) weights feed the pack through ordinary autograd; a second backward
    accumulates into the existing .grad like any other op."""
    R, U, D = 4, 6, 8
    g = torch.Generator().manual_seed(7)
    base = [torch.randn(U, D, generator=g, requires_grad=True) for _ in range(R)]
    bias = [torch.randn(U, generator=g, requires_grad=True) for _ in range(R)]
    base2 = [w.detach().clone().requires_grad_(True) for w in base]
    bias2 = [b.detach().clone().requires_grad_(True) for b in bias]

    def cum(ts):
        out, acc = [], None
        for t in ts:
            acc = t if acc is None else acc + t
            out.append(acc)
        return out

    gout = torch.randn(U, R * D + R, generator=g)
    for _ in range(2):
        pack_w_ext(cum(base), cum(bias)).backward(gout)
        torch.cat(cum(base2) + [torch.stack(cum(bias2), dim=1)], dim=1).backward(gout)
    for p, q in zip(base + bias, base2 + bias2):
        torch.testing.assert_close(p.grad, q.grad, rtol=0, atol=1e-6)
