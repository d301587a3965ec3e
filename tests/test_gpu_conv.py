"""Kernel-level parity of the convolutional stages either side of the hot path: the nodes of
adaptersis_b200/conv.py against the ATen operators the reference modules are made of (backbones/encoders.py:9-74,
backbones/decoders.py:92-164), forward and every gradient, in both precision modes.  Tolerances: fp32 mode 1e-4 on
activations / input gradients (weight gradients 1e-3 of their maximum: 1e5-term sums in another order), bf16 mode 2e-2."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2)


@pytest.mark.parametrize("mode,tol,wtol", [("fp32", 1e-4, 1e-3), ("bf16", 2e-2, 2e-2)])
@pytest.mark.parametrize("B,Cin,Cout,H,W,stride,pad", [(2, 3, 64, 30, 30, 2, 1), (2, 64, 128, 37, 37, 2, 0),
                                                       (1, 64, 64, 20, 23, 1, 1), (2, 128, 64, 9, 9, 2, 1)])
def test_conv2d_im2col_path(mode, tol, wtol, B, Cin, Cout, H, W, stride, pad):
    import adaptersis_b200 as asis
    from adaptersis_b200 import conv as Cv
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    want_gx = Cin % 8 == 0        # (the image itself never needs a gradient: the 3-channel col2im is not provided)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yr = F.conv2d(xr, wr, br, stride=stride, padding=pad)
    gy = torch.randn(yr.shape, generator=g).to(DEV)
    gr = torch.autograd.grad(yr, (xr, wr, br), gy)
    xo, wo, bo = _nhwc(x).requires_grad_(want_gx), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    with asis.precision(mode):
        y = Cv.conv2d(xo, wo, bo, stride, pad)
        go = torch.autograd.grad(y, ((xo,) if want_gx else ()) + (wo, bo), _nhwc(gy).to(y.dtype))
    assert relerr(_nchw(y).float(), yr) < tol
    if want_gx:
        assert relerr(_nchw(go[0]).float(), gr[0]) < tol
    assert relerr(go[-2].float(), gr[1]) < wtol and relerr(go[-1].float(), gr[2]) < wtol


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 64, 64, 30, 30), (1, 128, 64, 17, 21), (2, 192, 128, 12, 12), (1, 64, 256, 40, 9),
                                            (3, 64, 64, 1, 1)])
def test_conv3x3_implicit_gemm(B, Cin, Cout, H, W):
    """asis_conv3x3s1_gemm: forward, input gradient and weight gradient of a 3x3 / stride 1 / pad 1 convolution over
    zero-padded channels-last bf16 maps, against F.conv2d on the same bf16-rounded operands (fp32 arithmetic)."""
    from adaptersis_b200 import conv as Cv
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, Cin, H, W, generator=g).to(DEV).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).to(DEV).bfloat16().float()
    b = torch.randn(Cout, generator=g).to(DEV)
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, padding=1)
    gy = torch.randn(yr.shape, generator=g).to(DEV).bfloat16()
    gr = torch.autograd.grad(yr, (xr, wr, br), gy.float())
    xp = F.pad(_nhwc(x), (0, 0, 1, 1, 1, 1)).contiguous().requires_grad_(True)
    wo, bo = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yp = Cv.conv3x3_padded(xp, wo, bo)
    assert yp.shape == (B, H + 2, W + 2, Cout) and yp.dtype == torch.bfloat16
    gyp = F.pad(_nhwc(gy), (0, 0, 1, 1, 1, 1)).contiguous()          # the backward contract: zero border
    go = torch.autograd.grad(yp, (xp, wo, bo), gyp)
    y = yp[:, 1:-1, 1:-1]
    assert relerr(_nchw(y).float(), yr) < 8e-3                       # bf16 rounding of the output only
    assert relerr(_nchw(go[0][:, 1:-1, 1:-1]).float(), gr[0]) < 8e-3
    assert relerr(go[1], gr[1]) < 2e-3 and relerr(go[2], gr[2]) < 2e-3      # f32 outputs of bf16 products


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("pad_in,pad_out", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_batch_norm_node(mode, tol, relu, pad_in, pad_out):
    """nn.BatchNorm2d in training mode (+ ReLU) incl. running statistics, with every storage-padding combination."""
    from adaptersis_b200 import conv as Cv
    g = torch.Generator().manual_seed(3)
    B, C, H, W = 3, 64, 13, 11
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    x = (torch.randn(B, C, H, W, generator=g) * 2 + 3).to(DEV).to(dt)
    bn_r = nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn_r.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn_r.bias.copy_(torch.randn(C, generator=g))
    bn_o = nn.BatchNorm2d(C).to(DEV)
    bn_o.load_state_dict(bn_r.state_dict())
    xr = x.float().requires_grad_(True)
    yr = bn_r(xr)
    yr = F.relu(yr) if relu else yr
    gy = torch.randn(yr.shape, generator=g).to(DEV).to(dt)
    gr = torch.autograd.grad(yr, (xr, bn_r.weight, bn_r.bias), gy.float())
    junk = 7.0      # border values of an input in padded storage are not meaningful: they must not be read
    xs = F.pad(_nhwc(x), (0, 0, pad_in, pad_in, pad_in, pad_in), value=junk).contiguous().requires_grad_(True)
    y = Cv.batch_norm(xs, bn_o, relu=relu, pad_in=pad_in, pad_out=pad_out)
    assert y.shape == (B, H + 2 * pad_out, W + 2 * pad_out, C)
    gys = F.pad(_nhwc(gy), (0, 0, pad_out, pad_out, pad_out, pad_out), value=junk).contiguous()
    go = torch.autograd.grad(y, (xs, bn_o.weight, bn_o.bias), gys)
    crop = (lambda t, p: t[:, p:t.shape[1] - p, p:t.shape[2] - p])
    assert relerr(_nchw(crop(y, pad_out)).float(), yr) < tol
    if pad_out:      # zero border on the way out (the implicit GEMM reads it as the convolution's zero padding)
        assert float(y[:, 0].abs().max()) == 0 and float(y[:, :, -1].abs().max()) == 0
    assert relerr(_nchw(crop(go[0], pad_in)).float(), gr[0]) < tol
    if pad_in:
        assert float(go[0][:, 0].abs().max()) == 0 and float(go[0][:, :, 0].abs().max()) == 0
    assert relerr(go[1], gr[1]) < tol and relerr(go[2], gr[2]) < tol
    assert relerr(bn_o.running_mean, bn_r.running_mean) < tol and relerr(bn_o.running_var, bn_r.running_var) < tol


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H,W", [(30, 30), (13, 7)])
def test_maxpool_node(dt, H, W):
    from adaptersis_b200 import conv as Cv
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 64, H, W, generator=g).to(DEV).to(dt)
    x = torch.relu(x)                       # many exact ties at zero, as behind the stem's ReLU: the tie rule matters
    xr = x.float().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    gy = torch.randn(yr.shape, generator=g).to(DEV).to(dt)
    (gr,) = torch.autograd.grad(yr, xr, gy.float())
    xo = _nhwc(x).requires_grad_(True)
    y = Cv.maxpool3x3s2(xo)
    (go,) = torch.autograd.grad(y, xo, _nhwc(gy))
    assert torch.equal(_nchw(y).float(), yr)
    assert relerr(_nchw(go).float(), gr) < (1e-6 if dt == torch.float32 else 1e-2)


@pytest.mark.parametrize("dt,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("B,C,H,W,CO", [(2, 64, 21, 21, 2), (1, 64, 9, 14, 4), (1, 128, 5, 5, 1)])
def test_seg_head_node(dt, tol, B, C, H, W, CO):
    """decoders.py:125-129: nn.Upsample(x2, bilinear, align_corners=True) -> 3x3 / pad 1 convolution, as one node."""
    from adaptersis_b200 import conv as Cv
    g = torch.Generator().manual_seed(5)
    z = torch.randn(B, C, H, W, generator=g).to(DEV).to(dt)
    w = (torch.randn(CO, C, 3, 3, generator=g) * 0.1).to(DEV)
    b = torch.randn(CO, generator=g).to(DEV)
    zr, wr, br = z.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv2d(F.interpolate(zr, scale_factor=2, mode="bilinear", align_corners=True), wr, br, padding=1)
    gy = torch.randn(yr.shape, generator=g).to(DEV)
    gr = torch.autograd.grad(yr, (zr, wr, br), gy)
    zo, wo, bo = _nhwc(z).requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = Cv.seg_head(zo, wo, bo)
    go = torch.autograd.grad(y, (zo, wo, bo), _nhwc(gy))
    assert y.dtype == torch.float32 and relerr(_nchw(y), yr) < tol
    assert relerr(_nchw(go[0]).float(), gr[0]) < tol
    assert relerr(go[1], gr[1]) < max(tol, 1e-3) and relerr(go[2], gr[2]) < max(tol, 1e-3)


def _torch_spm(inplanes, embed_dim):
    """the reference layer stack (backbones/encoders.py:9-52) out of stock nn modules (BatchNorm2d = SyncBatchNorm on one rank)."""
    def cbr(cin, cout, stride, padding):
        return [nn.Conv2d(cin, cout, 3, stride, padding, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
    p = inplanes
    m = nn.Module()
    m.stem = nn.Sequential(*(cbr(3, p, 2, 1) + cbr(p, p, 1, 1) + cbr(p, p, 1, 1) + [nn.MaxPool2d(3, 2, 1)]))
    m.conv2 = nn.Sequential(*cbr(p, 2 * p, 2, 0))
    m.conv3 = nn.Sequential(*cbr(2 * p, 4 * p, 2, 0))
    m.conv4 = nn.Sequential(*cbr(4 * p, 8 * p, 2, 1))
    m.fc1, m.fc2, m.fc3, m.fc4 = (nn.Conv2d(p * k, embed_dim, 1) for k in (1, 2, 4, 8))
    return m


def _sd64(module):
    return {k: (v.detach().double().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in module.state_dict().items()}


def _grad_check(names, ours, sd, tol, skip=(), yard=None):
    """ours vs the fp64 oracle gradients in ``sd[n].grad``; ``yard`` (bf16 mode): the same oracle under
    torch.autocast(bfloat16) -- the reference's own arithmetic at that precision -- whose distance from fp64 sets the
    bound where it exceeds ``tol`` (3 x, the rule of tests/test_gpu_fullsize.py)."""
    bad = []
    for n, a in zip(names, ours):
        r = sd[n].grad
        if n.endswith(skip) or r is None or float(r.abs().max()) < 1e-12:
            continue
        e = relerr(a.float(), r)
        bound = tol if yard is None else max(tol, 3.0 * relerr(yard[n].grad.float(), r))
        if not e < bound:
            bad.append((n, e, bound))
    assert not bad, bad


def _sd32(module):
    return {k: (v.detach().float().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in module.state_dict().items()}


# Gradient parity of the two layer stacks.  ReLU' is discontinuous at 0: a pre-activation within rounding distance of
# zero takes the other branch in ANY two fp32 evaluations (tools/spm_debug.py: 1-3 such elements per layer for the stock
# fp32 modules and for ours alike, at different pixels), and on these small maps one flipped element moves a channel's
# gradients by 1-3 %.  So the fp64 oracle (oracle/encoder.py, the reference's layer stack restated) is evaluated twice:
# with its own ReLU for the forward outputs, and ON THE MASKS OF THE RUN UNDER TEST for the gradients -- the function is
# then smooth and the bound is the plain one: 1e-4 on outputs, 2e-3 on parameter gradients (1e5-term fp32 sums) in fp32
# mode; 2e-2 / 3e-2 in bf16 mode.  The spatial prior module keeps one more discontinuity the masks do not remove -- the
# arg-max of the stem's 3x3 max pooling between near-equal neighbours (~1 of 7e5 windows) -- and is six bf16 layers deep:
# 1e-2 on its gradients in fp32 mode (measured 6e-3 on stem.6.weight, everything else < 1e-3).  In bf16 mode the masks of
# a bf16 run differ from the fp64 ones at thousands of elements and BatchNorm over small maps amplifies bf16 rounding
# (PyTorch's own autocast gradients of the stem are 10-30 % from fp64 here): bound = max(tol, 3 x autocast's distance).
@pytest.mark.parametrize("mode,tol,gtol", [("fp32", 1e-4, 1e-2), ("bf16", 3e-2, 4e-2)])
def test_spatial_prior_module_vs_oracle(mode, tol, gtol):
    """FeatureEncoder end to end (stem, three stride-2 stages, 1x1 projections) at 292 x 292, batch 2."""
    import adaptersis_b200 as asis
    from adaptersis_b200.encoders import FeatureEncoder
    from oracle import encoder as o_enc
    torch.manual_seed(6)
    ours = FeatureEncoder(inplanes=64, embed_dim=128).to(DEV)
    with torch.no_grad():
        for n, p in ours.named_parameters():           # BatchNorm away from its (1, 0) initialisation
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    x = torch.rand(2, 3, 292, 292, device=DEV)
    names = [n for n, _ in ours.named_parameters() if not n.startswith("fc1")]
    rec = {}
    with asis.precision(mode):
        _, o2, o3, o4 = ours(x, need_c1=False, record=rec)
        gs = [torch.randn_like(o) for o in (o2, o3, o4)]
        go = torch.autograd.grad([o2, o3, o4], [dict(ours.named_parameters())[n] for n in names], gs)
    assert sorted(rec) == ["conv2.1.", "conv3.1.", "conv4.1.", "stem.1.", "stem.4.", "stem.7."]
    sd = _sd64(ours)
    with torch.no_grad():
        fwd = o_enc.spm(sd, x.double())[1:]
    for a, b in zip((o2, o3, o4), fwd):
        assert a.shape == b.shape and relerr(a.float(), b) < tol
    outs = o_enc.spm(sd, x.double(), relu_masks=rec)[1:]
    torch.autograd.backward(outs, [g.double() for g in gs])
    yard = None
    if mode == "bf16":
        yard = _sd32(ours)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            youts = o_enc.spm(yard, x)[1:]
        torch.autograd.backward(youts, [g.to(o.dtype) for g, o in zip(gs, youts)])
    _grad_check(names, go, sd, gtol, yard=yard)


@pytest.mark.parametrize("mode,tol,gtol", [("fp32", 1e-4, 2e-3), ("bf16", 2e-2, 3e-2)])
def test_feature_decoder_vs_oracle(mode, tol, gtol):
    """FeatureDecoder (decoders.py:92-164) at a 10 x 10 token grid: logits, the input gradient, every parameter gradient."""
    import adaptersis_b200 as asis
    from adaptersis_b200.decoders import FeatureDecoder
    from oracle import encoder as o_enc
    torch.manual_seed(7)
    feats = [64, 128, 64, 64, 64]
    ours = FeatureDecoder(img_size=140, embed_dim=64, num_classes=2, features=feats).to(DEV)
    with torch.no_grad():
        for n, p in ours.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(2, 3 * feats[0], 10, 10, device=DEV)
    gy = torch.randn(2, 2, 160, 160, device=DEV)
    names = [n for n, _ in ours.named_parameters()]
    rec = {}
    xo = x.clone().requires_grad_(True)
    with asis.precision(mode):
        y = ours(xo, record=rec)
        go = torch.autograd.grad(y, [xo] + [dict(ours.named_parameters())[n] for n in names], gy)
    sd = _sd64(ours)
    with torch.no_grad():
        yr = o_enc.feature_decoder(sd, x.double())
    assert y.shape == yr.shape and relerr(y.float(), yr) < tol
    sd["input"] = x.double().requires_grad_(True)
    o_enc.feature_decoder(sd, sd["input"], relu_masks=rec).backward(gy.double())
    yard = None
    if mode == "bf16":
        yard = _sd32(ours)
        yard["input"] = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yo = o_enc.feature_decoder(yard, yard["input"])
        yo.backward(gy.to(yo.dtype))
    # (a convolution bias in front of BatchNorm has an analytically zero gradient: rounding noise on both sides)
    _grad_check(["input"] + names, go, sd, gtol, skip=(".0.bias",), yard=yard)
