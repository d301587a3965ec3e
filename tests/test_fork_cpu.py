"""Host-side logic of Fn.layer_norm_fork (CPU, the two kernels replaced by torch stand-ins): the node must hand the
pass-through gradient to the LayerNorm backward as its residual operand and produce the same gradients as the two
separate uses it replaces (adapter_blocks.py:127-143: a tensor feeds a LayerNorm and continues as the residual)."""
import torch
import torch.nn.functional as F

from adaptersis_b200 import functional as Fn
from adaptersis_b200 import kernels as K


def _fake_fwd(x2, w, b, eps, out_dtype):
    mean = x2.float().mean(1)
    rstd = (x2.float().var(1, unbiased=False) + eps).rsqrt()
    y = ((x2.float() - mean[:, None]) * rstd[:, None] * w + b).to(out_dtype)
    return y, mean, rstd


def _fake_bwd(dy2, x2, w, mean, rstd, dres=None, want_param_grads=True):
    calls.append(dres is not None)
    xh = (x2.float() - mean[:, None]) * rstd[:, None]
    g = dy2.float() * w
    dx = rstd[:, None] * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True))
    if dres is not None:
        dx = dx + dres.float()
    return dx, (dy2.float() * xh).sum(0), dy2.float().sum(0)


calls = []


def test_fork_matches_two_separate_uses(monkeypatch):
    monkeypatch.setattr(K, "layernorm_forward", _fake_fwd)
    monkeypatch.setattr(K, "layernorm_backward", _fake_bwd)
    torch.manual_seed(0)
    x = torch.randn(2, 5, 16, requires_grad=True)
    w = torch.randn(16, requires_grad=True)
    b = torch.randn(16, requires_grad=True)
    a, c = torch.randn(2, 5, 16), torch.randn(2, 5, 16)
    with Fn.precision("fp32"):
        y, xp = Fn.layer_norm_fork(x, w, b, 1e-6)
        assert xp.data_ptr() == x.data_ptr() and xp is not x and torch.equal(xp, x)
        z, xq = Fn.layer_norm_fork(xp, w, b, 1e-6)          # chained forks: three uses of x, no gradient add outside
        ((y * a).sum() + (z * z).sum() + (xq * c).sum()).backward()
    got = (x.grad.clone(), w.grad.clone(), b.grad.clone())
    assert calls == [True, True]
    x.grad = w.grad = b.grad = None
    y = F.layer_norm(x, (16,), w, b, 1e-6)
    ((y * a).sum() + (y * y).sum() + (x * c).sum()).backward()
    for g, r in zip(got, (x.grad, w.grad, b.grad)):
        assert torch.allclose(g, r, rtol=1e-4, atol=1e-5), float((g - r).abs().max())
    # unused pass-through output: plain LayerNorm backward; no-grad input: no node at all
    calls.clear()
    x.grad = None
    with Fn.precision("fp32"):
        y, _ = Fn.layer_norm_fork(x, w, b, 1e-6)
        y.sum().backward()
    assert calls == [False]
    with Fn.precision("fp32"), torch.no_grad():
        y, xp = Fn.layer_norm_fork(x, w, b, 1e-6)
    assert xp is x and not y.requires_grad
