"""GPU parity of the dense / normalisation kernels in fp32 parity mode (FFMA kernels) vs plain
torch fp32 on the same inputs (tolerance 1e-4 max-relative), through the C ABI."""
import pytest
import torch

from conftest import relerr
import adaptersis_b200 as asis
from adaptersis_b200 import kernels as K
from adaptersis_b200._lib import (EPI_ACCUMULATE, EPI_DGELU, EPI_GELU, EPI_NONE, EPI_SCALE_RESIDUAL, F32, MAJOR_K,
                                  MAJOR_MN)
from oracle import layers as o_layers

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


@pytest.mark.parametrize("R,C", [(37, 64), (1765, 1024), (100, 384), (64, 768), (9, 32)])
def test_layernorm(R, C):
    torch.manual_seed(0)
    x = torch.randn(R, C, device=DEV) * 2 + 0.5
    w = torch.randn(C, device=DEV)
    b = torch.randn(C, device=DEV)
    dy = torch.randn(R, C, device=DEV)
    dres = torch.randn(R, C, device=DEV)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-6)
    gx, gw, gb = torch.autograd.grad(yr, (xr, wr, br), dy)
    y, mean, rstd = K.layernorm_forward(x, w, b, 1e-6, torch.float32)
    assert relerr(y, yr) < TOL
    dx, dw, db = K.layernorm_backward(dy, x, w, mean, rstd, dres)
    assert relerr(dx, gx + dres) < TOL and relerr(dw, gw) < TOL and relerr(db, gb) < TOL
    yb, _, _ = K.layernorm_forward(x.bfloat16(), w, b, 1e-6, torch.bfloat16)
    assert relerr(yb.float(), torch.nn.functional.layer_norm(x.bfloat16().float(), (C,), w, b, 1e-6)) < 2e-2


@pytest.mark.parametrize("M,N,K_", [(70, 96, 40), (257, 192, 64), (128, 64, 1024), (5, 8, 3)])
def test_gemm_f32_majors_and_epilogues(M, N, K_):
    torch.manual_seed(1)
    A = torch.randn(M, K_, device=DEV)
    B = torch.randn(N, K_, device=DEV)
    ref = A @ B.t()
    for am in (MAJOR_K, MAJOR_MN):
        for bm in (MAJOR_K, MAJOR_MN):
            a = A if am == MAJOR_K else A.t().contiguous()
            b = B if bm == MAJOR_K else B.t().contiguous()
            c, _ = K.gemm(F32, a, am, b, bm, M, N, K_, torch.float32)
            assert relerr(c, ref) < TOL
    bias = torch.randn(N, device=DEV)
    gamma = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    c, _ = K.gemm(F32, A, MAJOR_K, B, MAJOR_K, M, N, K_, torch.float32, bias=bias)
    assert relerr(c, ref + bias) < TOL
    c, aux = K.gemm(F32, A, MAJOR_K, B, MAJOR_K, M, N, K_, torch.float32, epilogue=EPI_GELU, bias=bias,
                    want_aux_dtype=torch.float32)
    assert relerr(aux, ref + bias) < TOL and relerr(c, torch.nn.functional.gelu(ref + bias)) < TOL
    c, aux = K.gemm(F32, A, MAJOR_K, B, MAJOR_K, M, N, K_, torch.float32, epilogue=EPI_SCALE_RESIDUAL, bias=bias,
                    gamma=gamma, residual=res, want_aux_dtype=torch.float32)
    assert relerr(c, res + gamma * (ref + bias)) < TOL and relerr(aux, ref + bias) < TOL
    h = torch.randn(M, N, device=DEV, requires_grad=True)
    (dg,) = torch.autograd.grad(torch.nn.functional.gelu(h), h, torch.ones_like(h))
    c, _ = K.gemm(F32, A, MAJOR_K, B, MAJOR_K, M, N, K_, torch.float32, epilogue=EPI_DGELU, aux=h.detach())
    assert relerr(c, ref * dg) < TOL
    acc = torch.randn(M, N, device=DEV)
    want = acc + ref
    c, _ = K.gemm(F32, A, MAJOR_K, B, MAJOR_K, M, N, K_, torch.float32, epilogue=EPI_ACCUMULATE, out=acc)
    assert relerr(c, want) < TOL


def test_colsum_scale_add_cast():
    torch.manual_seed(2)
    X = torch.randn(1000, 192, device=DEV)
    Y = torch.randn(1000, 192, device=DEV)
    assert relerr(K.colsum(X), X.sum(0)) < TOL
    assert relerr(K.colsum(X, Y), (X * Y).sum(0)) < TOL
    assert relerr(K.colsum(X.bfloat16()), X.bfloat16().float().sum(0)) < TOL
    g = torch.randn(192, device=DEV)
    assert relerr(K.scale_cols(X, g, torch.float32), X * g) < 1e-6
    assert relerr(K.add(X, Y.bfloat16(), torch.float32), X + Y.bfloat16().float()) < 1e-6
    assert torch.equal(K.cast(X, torch.bfloat16), X.bfloat16())
    v = torch.randn(1003, device=DEV)
    assert torch.equal(K.cast(v[:1001].clone(), torch.bfloat16), v[:1001].bfloat16())


@pytest.mark.parametrize("B,T,H,hd", [(2, 19, 4, 16), (1, 300, 3, 64), (2, 77, 2, 32)])
def test_attention_f32(B, T, H, hd):
    torch.manual_seed(3)
    C = H * hd
    qkv = torch.randn(B, T, 3 * C, device=DEV, requires_grad=True)
    q, k, v = qkv.view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2), -1)
    ref = (p @ v).transpose(1, 2).reshape(B, T, C)
    dout = torch.randn(B, T, C, device=DEV)
    (gref,) = torch.autograd.grad(ref, qkv, dout)
    with asis.precision("fp32"):
        x = qkv.detach().clone().requires_grad_(True)
        out = asis.functional.attention(x, H)
        assert relerr(out, ref) < TOL
        (g,) = torch.autograd.grad(out, x, dout)
        assert relerr(g, gref) < TOL


def test_patchify_and_dwconv():
    torch.manual_seed(4)
    img = torch.rand(2, 3, 42, 28, device=DEV)
    cols = K.patchify(img, 14, torch.float32, 592)
    ref = img.view(2, 3, 3, 14, 2, 14).permute(0, 2, 4, 1, 3, 5).reshape(12, 588)
    assert torch.equal(cols[:, :588], ref) and float(cols[:, 588:].abs().max()) == 0.0
    with pytest.raises(AssertionError, match="multiple of patch"):
        K.patchify(torch.rand(1, 3, 40, 28, device=DEV), 14, torch.float32, 592)
    C = 64
    maps = [(7, 5), (4, 3)]
    ntok = sum(h * w for h, w in maps)
    for fuse in (False, True):
        x = torch.randn(2, ntok, C, device=DEV, requires_grad=True)
        w = torch.randn(C, 1, 3, 3, device=DEV, requires_grad=True)
        b = torch.randn(C, device=DEV, requires_grad=True)
        outs, s = [], 0
        for hh, ww in maps:
            t = x[:, s:s + hh * ww].transpose(1, 2).reshape(2, C, hh, ww)
            t = torch.nn.functional.conv2d(t, w, b, padding=1, groups=C)
            outs.append(t.flatten(2).transpose(1, 2))
            s += hh * ww
        ref = torch.cat(outs, 1)
        if fuse:
            ref = torch.nn.functional.gelu(ref)
        dy = torch.randn_like(ref)
        gx, gw, gb = torch.autograd.grad(ref, (x, w, b), dy)
        x2 = x.detach().clone().requires_grad_(True)
        w2 = w.detach().clone().requires_grad_(True)
        b2 = b.detach().clone().requires_grad_(True)
        y = asis.functional.DWConvFunction.apply(x2, w2, b2, maps, fuse)
        assert relerr(y, ref) < TOL
        hx, hw, hb = torch.autograd.grad(y, (x2, w2, b2), dy)
        assert relerr(hx, gx) < TOL and relerr(hw, gw) < TOL and relerr(hb, gb) < TOL


@pytest.mark.parametrize("B,C,H,W,dtype", [(2, 64, 21, 21, torch.float32), (1, 16, 7, 5, torch.float32),
                                           (2, 64, 42, 42, torch.bfloat16), (1, 8, 3, 2, torch.bfloat16)])
def test_upsample2x_matches_torch(B, C, H, W, dtype):
    # FeatureDecoder's nn.Upsample(scale_factor=2, bilinear, align_corners=True): forward and backward
    from adaptersis_b200 import functional as Fn
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, H, W, generator=g).to(DEV, dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gy = torch.randn(B, C, 2 * H, 2 * W, generator=g).to(DEV, dtype)
    y = Fn.upsample2x(x)
    (gx,) = torch.autograd.grad(y, x, gy)
    xr = x.detach().float().requires_grad_(True)
    yr = torch.nn.functional.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    (gxr,) = torch.autograd.grad(yr, xr, gy.float())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert y.shape == yr.shape and relerr(y.float(), yr) < tol
    assert relerr(gx.float(), gxr) < tol


def test_frames_to_batch_matches_host_pipeline():
    # tools/dataset.py:111-118: torch.from_numpy(img_np.transpose(2, 0, 1)) / 255.0 and mask.long(), bit for bit
    g = torch.Generator().manual_seed(9)
    frames = torch.randint(0, 256, (3, 70, 42, 3), generator=g, dtype=torch.uint8)
    masks = torch.randint(0, 2, (3, 70, 42), generator=g, dtype=torch.uint8)
    ref_img = frames.permute(0, 3, 1, 2) / 255.0
    ref_tgt = masks.long()
    img, tgt = K.frames_to_batch(frames.to(DEV), masks.to(DEV))
    assert img.dtype == torch.float32 and tgt.dtype == torch.int64
    assert torch.equal(img.cpu(), ref_img) and torch.equal(tgt.cpu(), ref_tgt)
    img2, none = K.frames_to_batch(frames.to(DEV))
    assert none is None and torch.equal(img2.cpu(), ref_img)


def test_fused_sgd_matches_torch():
    # train.py:178-189: SGD(momentum=0.99, weight_decay=3e-5); three steps incl. the first (buffer creation), odd sizes,
    # a parameter without gradient, 60 tensors (two launches of the 48-entry pointer table)
    from adaptersis_b200.trainer import FusedSGD
    g = torch.Generator().manual_seed(11)
    shapes = [(1024, 1024), (7,), (3, 5, 2), (4096,), (1,), (129, 33)] * 10
    pa = [torch.randn(s, generator=g).to(DEV).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedSGD(pa, lr=0.01, momentum=0.99, weight_decay=3e-5)
    ob = torch.optim.SGD(pb, lr=0.01, momentum=0.99, weight_decay=3e-5, foreach=False, fused=False)
    for step in range(3):
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == 4 and step == 0:
                a.grad = b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
        if step == 0:
            for grp in oa.param_groups:
                grp["lr"] = 0.005          # a scheduler changes the learning rate between steps
            for grp in ob.param_groups:
                grp["lr"] = 0.005
    for a, b in zip(pa, pb):
        assert relerr(a, b) < 1e-6
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["state"].keys() == sb["state"].keys()
    for k in sa["state"]:
        assert relerr(sa["state"][k]["momentum_buffer"], sb["state"][k]["momentum_buffer"]) < 1e-6
