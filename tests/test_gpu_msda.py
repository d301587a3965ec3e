"""GPU parity: multi-scale deformable attention kernels (through the C ABI) vs the oracle and the
reference-generated golden vectors.  fp32 tolerance 1e-4 max-relative (north_star), bf16 2e-2."""
import pytest
import torch

from conftest import relerr
import adaptersis_b200 as asis
from adaptersis_b200 import kernels as K
from oracle import msda as o_msda

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL32 = 1e-4
TOLBF = 2e-2


def _lsi(ss):
    return torch.cat([ss.new_zeros(1), ss.prod(1).cumsum(0)[:-1]])


def _rand_case(seed, N, Lq, M, D, shapes, P, spread=4.0, oob=0.05):
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g)
    base = torch.rand(N, Lq, 1, 1, 1, 2, generator=g)
    loc = base.expand(N, Lq, M, L, P, 2).clone()
    for l, (H, W) in enumerate(shapes):
        loc[:, :, :, l] += (torch.rand(N, Lq, M, P, 2, generator=g) * 2 - 1) * spread / torch.tensor([W, H], dtype=torch.float32)
    far = torch.rand(N, Lq, M, L, P, generator=g) < oob
    loc[far] = torch.rand(int(far.sum()), 2, generator=g) * 1.2 - 0.1
    # d(out)/d(loc) is discontinuous where a pixel coordinate is an integer; two correct fp32
    # evaluations may floor() differently there, so keep samples 1e-3 px away from cell borders
    for l, (H, W) in enumerate(shapes):
        size = torch.tensor([W, H], dtype=torch.float32)
        pix = loc[:, :, :, l] * size - 0.5
        frac = pix - pix.floor()
        near = (frac < 1e-3) | (frac > 1 - 1e-3)
        loc[:, :, :, l] = torch.where(near, loc[:, :, :, l] + 3e-3 / size, loc[:, :, :, l])
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, generator=g)
    return value, torch.as_tensor(shapes, dtype=torch.long), loc, aw, gout


def test_golden_forward_backward(golden):
    for case in golden("msda_core.pt"):
        ss = case["spatial_shapes"].to(DEV)
        lsi = case["level_start_index"].to(DEV)
        v = case["value"].to(DEV).requires_grad_(True)
        loc = case["loc"].to(DEV).requires_grad_(True)
        aw = case["aw"].to(DEV).requires_grad_(True)
        out = asis.MSDeformAttnFunction.apply(v, ss, lsi, loc, aw, 64)
        assert relerr(out, case["out"]) < TOL32
        gv, gl, ga = torch.autograd.grad(out, (v, loc, aw), case["grad_out"].to(DEV))
        assert relerr(gv, case["grad_value"]) < TOL32
        assert relerr(gl, case["grad_loc"]) < TOL32
        assert relerr(ga, case["grad_aw"]) < TOL32


@pytest.mark.parametrize("name,N,Lq,M,D,shapes,P", [
    ("injector", 2, 1764, 8, 128, [(73, 73), (36, 36), (18, 18)], 4),
    ("extractor", 2, 6949, 8, 128, [(42, 42)], 4),
    ("m16", 1, 1024, 16, 64, [(72, 72), (36, 36), (18, 18)], 4),
    ("d96_p3", 1, 333, 8, 96, [(20, 31), (9, 7)], 3),
    ("d32", 2, 500, 4, 32, [(16, 16), (8, 8)], 4),
    ("d8", 1, 77, 3, 8, [(5, 6)], 2),
])
def test_against_oracle_fp32(name, N, Lq, M, D, shapes, P):
    value, ss, loc, aw, gout = _rand_case(3, N, Lq, M, D, shapes, P)
    vo = value.clone().requires_grad_(True)
    lo = loc.clone().requires_grad_(True)
    ao = aw.clone().requires_grad_(True)
    ref = o_msda.msda_core(vo, shapes, lo, ao)
    rgv, rgl, rga = torch.autograd.grad(ref, (vo, lo, ao), gout)
    vd, ld, ad = value.to(DEV), loc.to(DEV), aw.to(DEV)
    ssd = ss.to(DEV)
    out = K.msda_forward(vd, ssd, _lsi(ssd), ld, ad)
    assert relerr(out, ref) < TOL32
    gv, gl, ga = K.msda_backward(vd, ssd, _lsi(ssd), ld, ad, gout.to(DEV))
    assert relerr(gv, rgv) < TOL32
    assert relerr(gl, rgl) < TOL32
    assert relerr(ga, rga) < TOL32
    # run-to-run determinism of the atomic-free backward: bit-identical
    gv2, gl2, ga2 = K.msda_backward(vd, ssd, _lsi(ssd), ld, ad, gout.to(DEV))
    assert torch.equal(gv, gv2) and torch.equal(gl, gl2) and torch.equal(ga, ga2)


def test_bf16_value_variant():
    shapes = [(73, 73), (36, 36), (18, 18)]
    value, ss, loc, aw, gout = _rand_case(5, 2, 1764, 8, 128, shapes, 4)
    vb = value.bfloat16()
    gb = gout.bfloat16()
    vo = vb.float().requires_grad_(True)
    lo = loc.clone().requires_grad_(True)
    ao = aw.clone().requires_grad_(True)
    ref = o_msda.msda_core(vo, shapes, lo, ao)
    rgv, rgl, rga = torch.autograd.grad(ref, (vo, lo, ao), gb.float())
    ssd = ss.to(DEV)
    out = K.msda_forward(vb.to(DEV), ssd, _lsi(ssd), loc.to(DEV), aw.to(DEV))
    assert out.dtype == torch.bfloat16 and relerr(out.float(), ref) < TOLBF
    gv, gl, ga = K.msda_backward(vb.to(DEV), ssd, _lsi(ssd), loc.to(DEV), aw.to(DEV), gb.to(DEV))
    assert relerr(gv.float(), rgv) < TOLBF and relerr(gl, rgl) < TOLBF and relerr(ga, rga) < TOLBF


def test_known_answers():
    # (ii) cell-centre sampling with one-hot weights reproduces value rows exactly
    H, W, M, D = 6, 7, 2, 16
    value = torch.randn(1, H * W, M, D, device=DEV)
    ss = torch.tensor([[H, W]], device=DEV)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    cx = ((xs.reshape(-1).float() + 0.5) / W).to(DEV)
    cy = ((ys.reshape(-1).float() + 0.5) / H).to(DEV)
    loc = torch.stack([cx, cy], -1).view(1, H * W, 1, 1, 1, 2).expand(1, H * W, M, 1, 2, 2).contiguous()
    aw = torch.zeros(1, H * W, M, 1, 2, device=DEV)
    aw[..., 0] = 1.0
    out = K.msda_forward(value, ss, _lsi(ss), loc, aw)
    assert relerr(out.view(1, H * W, M, D), value) < 1e-6
    # (iii) locations more than a pixel outside the map give exact zeros
    far = loc.clone()
    far[..., 0] = -0.5
    assert float(K.msda_forward(value, ss, _lsi(ss), far, aw).abs().max()) == 0.0
    far[..., 0] = 1.0 + 1.5 / W
    assert float(K.msda_forward(value, ss, _lsi(ss), far, aw).abs().max()) == 0.0
    gv, gl, ga = K.msda_backward(value, ss, _lsi(ss), far, aw, torch.ones_like(out))
    assert float(gv.abs().max()) == 0.0 and float(gl.abs().max()) == 0.0 and float(ga.abs().max()) == 0.0


def test_long_bucket_path():
    # every query samples the same pixel: one bucket receives Lq*P*4 contributions (slow path)
    H, W, M, D, Lq, P = 4, 4, 1, 16, 300, 2
    g = torch.Generator().manual_seed(1)
    value = torch.randn(1, H * W, M, D, generator=g)
    loc = torch.full((1, Lq, M, 1, P, 2), 0.4) + 0.01 * torch.rand(1, Lq, M, 1, P, 2, generator=g)
    aw = torch.rand(1, Lq, M, 1, P, generator=g)
    gout = torch.randn(1, Lq, M * D, generator=g)
    vo = value.clone().requires_grad_(True)
    ref = o_msda.msda_core(vo, [(H, W)], loc, aw)
    (rgv,) = torch.autograd.grad(ref, (vo,), gout)
    ss = torch.tensor([[H, W]], device=DEV)
    gv, _, _ = K.msda_backward(value.to(DEV), ss, _lsi(ss), loc.to(DEV), aw.to(DEV), gout.to(DEV))
    assert relerr(gv, rgv) < TOL32


@pytest.mark.parametrize("name,N,Lq,M,D,shapes,P", [
    # a level with more pixels than the fill kernel keeps cursors for in shared memory (16-bit cursors:
    # 81920 pixels) -> processed window by window
    ("fill_windows", 1, 3000, 2, 8, [(300, 300)], 4),
    # cursor table capped at 256 MB -> one query chunk of 17000 queries -> a bucket may exceed 65535
    # entries -> 32-bit cursors (40960 pixels per window) -- both rarely used paths of the backward
    ("fill_int_cursors", 12, 17000, 16, 8, [(424, 424)], 4),
])
def test_backward_large_value_maps(name, N, Lq, M, D, shapes, P):
    g = torch.Generator().manual_seed(21)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g).to(DEV)
    loc = (torch.rand(N, Lq, M, len(shapes), P, 2, generator=g) * 1.04 - 0.02).to(DEV)
    # half of the samples crowd into one corner of the map: long per-pixel lists
    loc[:, ::2] = loc[:, ::2] * 0.05
    aw = torch.softmax(torch.randn(N, Lq, M, len(shapes) * P, generator=g), -1).view(N, Lq, M, len(shapes), P).to(DEV)
    gout = torch.randn(N, Lq, M * D, generator=g).to(DEV)
    ss = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
    vo, lo, ao = value.clone().requires_grad_(True), loc.clone().requires_grad_(True), aw.clone().requires_grad_(True)
    ref = o_msda.msda_core(vo, shapes, lo, ao)            # the oracle's plain-PyTorch formulation, on the GPU
    rgv, rgl, rga = torch.autograd.grad(ref, (vo, lo, ao), gout)
    out = K.msda_forward(value, ss, _lsi(ss), loc, aw)
    assert relerr(out, ref) < TOL32
    gv, gl, ga = K.msda_backward(value, ss, _lsi(ss), loc, aw, gout)
    assert relerr(gv, rgv) < TOL32 and relerr(ga, rga) < TOL32
    gv2, _, _ = K.msda_backward(value, ss, _lsi(ss), loc, aw, gout)
    assert torch.equal(gv, gv2)


def test_module_golden(golden):
    g = golden("msda_module.pt")
    m = asis.MSDeformAttn(**g["cfg"]).to(DEV)
    m.load_state_dict(g["sd"])
    ss, lsi = g["spatial_shapes"].to(DEV), g["level_start_index"].to(DEV)
    with asis.precision("fp32"):
        out = m(g["query"].to(DEV), g["ref"].to(DEV), g["feat"].to(DEV), ss, lsi, None)
        assert relerr(out, g["out"]) < TOL32
        out = m(g["query"].to(DEV), g["ref"].to(DEV), g["feat"].to(DEV), ss, lsi, g["mask"].to(DEV))
        assert relerr(out, g["out_masked"]) < TOL32
        out = m(g["query"].to(DEV), g["ref4"].to(DEV), g["feat"].to(DEV), ss, lsi, None)
        assert relerr(out, g["out_box"]) < TOL32
        with pytest.raises(ValueError, match="2 or 4"):
            m(g["query"].to(DEV), torch.rand(2, 25, 3, 3, device=DEV), g["feat"].to(DEV), ss, lsi, None)
        with pytest.raises(AssertionError):
            m(g["query"].to(DEV), g["ref"].to(DEV), g["feat"].to(DEV)[:, :-1], ss.clone(), lsi, None)
