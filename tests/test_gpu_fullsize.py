"""Full-size parity of the BENCHMARKED object: `TrainStep` at BASELINE.json's dimensions -- ViT-L/14 (C=1024,
depth 24, 16 heads) with M=8 / D=128 adapters (config[2]) and ViT-B/14 (C=768, depth 12) with M=12 / D=64
adapters (config[1]) -- one 588x588 frame, stress-initialised so that every path carries signal, in BOTH
precision modes, against the oracle (`oracle.encoder`: train.py:275-428 restated) evaluated ON THE GPU in fp64
and in fp32 from the same parameters.

What is compared: encoder state (x, c), the 3C x 42 x 42 decoder input, the 588^2 logits, the loss, EVERY
parameter gradient of the step, and the full argmax mask.

Tolerances (north_star): fp32 mode 1e-4 on activations / logits; bf16 mode 2e-2.  Gradients: plain PyTorch
evaluating the same graph (the reference's arithmetic) is itself 1e-5..1e-2 away from the fp64 evaluation in fp32
(summation order; BatchNorm with batch statistics; `d out / d loc` of bilinear sampling is discontinuous at pixel
borders) and 5..40 % away under bf16 autocast (the dice loss of a softmax-of-softmax is very flat).  The CUDA path
is therefore required to be within the north_star figure OR no further from the fp64 truth than 4x (fp32 mode) / 3x
(bf16 mode) what PyTorch at the same precision is, per parameter -- the tolerance is set by data, not by hand.
The error table is written to gpurun_out/ and committed under profiles/.
"""
import json
import os

import pytest
import torch

from conftest import ROOT, relerr, relerr_rms
from synth import synth_image
import adaptersis_b200 as asis
from adaptersis_b200.trainer import TrainStep
from oracle import encoder as o_enc

pytestmark = pytest.mark.gpu
DEV = "cuda"

CONFIGS = {
    "vit_large": dict(arch="vit_large", adapter_heads=8),      # config[2]: train.py verbatim (M=8, D=128)
    "vit_base": dict(arch="vit_base", adapter_heads=12),       # config[1]: dim 768, num_heads 12 -> D=64
}


def stress_init(module, gen):
    """Same recipe as tests/golden/make_golden.py: non-zero injector gamma / LayerScale, learned offsets and
    attention weights, biases -- otherwise CAViT is an identity and the parity is vacuous.
    The two MSDA projections are scaled by 1/sqrt(fan_in) so that, at C = 768 / 1024, learned offsets stay at
    ~1.5 px and the attention logits at ~1 (the toy fixtures' 0.3 / 0.5 at C = 32 give the same magnitudes);
    unscaled they would be +-10 px and one-hot softmaxes: a chaotic map, not a trained adapter."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            r = torch.randn(p.shape, generator=gen).to(p.device)
            fan = p.shape[-1] ** 0.5 if p.dim() > 1 else 1.0
            if name.endswith("gamma"):
                p.copy_(0.5 + 0.2 * r)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * r)
            elif "sampling_offsets.weight" in name:
                p.copy_(1.5 / fan * r)
            elif "attention_weights.weight" in name:
                p.copy_(1.0 / fan * r)
            elif name.endswith("bias") and "sampling_offsets" not in name:
                p.copy_(0.05 * r)
            elif name.endswith(("cls_token", "mask_token")):
                p.copy_(0.1 * r)
    return module


def _named(ts):
    named = {}
    for tag, mod in (("vit", ts.encoder.model), ("spm", ts.encoder.backbone_encoder), ("inj", ts.encoder.cross_vit),
                     ("ext", ts.encoder.cross_cnn), ("dec", ts.seg_decoder)):
        for k, p in mod.named_parameters():
            named[f"{tag}.{k}"] = p
    return named


def _oracle(sds, heads, img, target, dtype, amp=False):
    """oracle forward + backward on the GPU in `dtype` (plain PyTorch ops; TF32 off).  ``amp``: the same graph under
    torch.autocast(bfloat16) -- what PyTorch's own mixed precision does to the reference's arithmetic."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        w = {tag: {k: (v.detach().to(dtype).requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                   for k, v in sd.items()} for tag, sd in sds.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            res = o_enc.adapter_encoder(w["vit"], w["spm"], w["inj"], w["ext"], img.to(dtype), heads)
            logits = o_enc.feature_decoder(w["dec"], res["feat"])
        logits = torch.nn.functional.interpolate(logits.float() if amp else logits, size=target.shape[-2:], mode="bilinear")
        loss = o_enc.dice_loss(torch.softmax(logits, 1), target)
        loss.backward()
        grads = {f"{tag}.{k}": v.grad for tag, sd in w.items() for k, v in sd.items()
                 if v.is_floating_point() and v.requires_grad and v.grad is not None}
        out = dict(x=res["x"].detach(), c=res["c"].detach(), feat=res["feat"].detach(), logits=logits.detach(),
                   loss=loss.detach(), grads=grads)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out


def _ours(ts, img, target):
    for p in ts.parameters():
        p.grad = None
    internals = {}
    loss = ts.forward_loss(img, target, internals)
    loss.backward()
    ts.reducer.finish()
    out = {k: v.detach().float() for k, v in internals.items()}
    out["grads"] = {k: p.grad.detach().float() for k, p in _named(ts).items() if p.grad is not None}
    return out


@pytest.fixture(scope="module", params=list(CONFIGS))
def fullsize(request):
    cfg = CONFIGS[request.param]
    gen = torch.Generator().manual_seed(2024)
    torch.manual_seed(2024)
    ts = TrainStep(arch=cfg["arch"], adapter_heads=cfg["adapter_heads"], device=DEV, precision="fp32")
    stress_init(ts, gen)
    ts.encoder.backbone_encoder.train()
    ts.seg_decoder.train()
    img = synth_image(gen, 1, 588).to(DEV)
    target = (torch.rand(1, 588, 588, generator=gen) < 0.4).long().to(DEV)
    sds = {"vit": ts.encoder.model.state_dict(), "spm": ts.encoder.backbone_encoder.state_dict(),
           "inj": ts.encoder.cross_vit.state_dict(), "ext": ts.encoder.cross_cnn.state_dict(),
           "dec": ts.seg_decoder.state_dict()}
    heads = ts.encoder.model.num_heads
    o64 = _oracle(sds, heads, img, target, torch.float64)
    o32 = _oracle(sds, heads, img, target, torch.float32)
    o16 = _oracle(sds, heads, img, target, torch.float32, amp=True)
    torch.cuda.empty_cache()
    return dict(name=request.param, ts=ts, img=img, target=target, o64=o64, o32=o32, o16=o16, report={})


def _mask_report(logits, o64, tol):
    ours = logits.argmax(1)
    ref = o64["logits"].argmax(1)
    flips = ours != ref
    margin = (o64["logits"][:, 1] - o64["logits"][:, 0]).abs()
    scale = float(o64["logits"].abs().max())
    near = margin <= 2 * tol * scale            # each logit may move by tol x max|logit|; the margin is a difference of two
    bad = flips & ~near
    return dict(pixels=int(flips.numel()), flips=int(flips.sum()), near_ties=int(near.sum()),
                flips_at_determined_pixels=int(bad.sum()),
                max_margin_at_flips=float(margin[flips].max()) if bool(flips.any()) else 0.0, logit_absmax=scale)


def _compare(fs, mode, act_tol, grad_floor, grad_factor):
    # the yardstick: plain PyTorch evaluating the same graph at the precision of `mode` (fp32 / bf16 autocast)
    ts, o64, o32 = fs["ts"], fs["o64"], fs["o32" if mode == "fp32" else "o16"]
    ts.precision = mode
    r = _ours(ts, fs["img"], fs["target"])
    rep = {"mode": mode, "activations": {}, "grads": {}}
    fails = []
    for k in ("x", "c", "feat", "logits"):
        rep["activations"][k] = dict(ours=relerr(r[k], o64[k]), ours_rms=relerr_rms(r[k], o64[k]),
                                     torch_same_precision=relerr(o32[k], o64[k]))
        # north_star's tolerance, or -- where plain-PyTorch fp32 (the reference's arithmetic) is itself further than that
        # from the fp64 evaluation -- twice the distance PyTorch fp32 has
        if not rep["activations"][k]["ours"] < max(act_tol, 2.0 * rep["activations"][k]["torch_same_precision"]):
            fails.append((k, rep["activations"][k]))
    rep["loss"] = dict(ours=float(r["loss"]), fp64=float(o64["loss"]), torch_same_precision=float(o32["loss"]))
    if not abs(float(r["loss"]) - float(o64["loss"])) < (1e-5 if mode == "fp32" else 2e-3):
        fails.append(("loss", rep["loss"]))
    rep["mask"] = _mask_report(r["logits"], o64, act_tol)
    # bit-exact wherever the decision is numerically determined at the tolerance granted to the logits
    if rep["mask"]["flips_at_determined_pixels"] != 0:
        fails.append(("mask", rep["mask"]))
    missing = [k for k in o64["grads"] if k not in r["grads"]]
    assert not missing, missing
    worst = []
    for k, g64 in o64["grads"].items():
        if float(g64.abs().max()) < 1e-12:          # analytically zero (conv bias in front of BatchNorm)
            continue
        e_ours = relerr(r["grads"][k], g64)
        e_t32 = relerr(o32["grads"][k], g64)
        rep["grads"][k] = dict(ours=e_ours, torch_same_precision=e_t32, numel=g64.numel())
        worst.append((e_ours, e_t32, k))
        # ReLU'(0) is a discontinuity: the convolutional stages (spm.*, dec.*) flip a handful of near-zero pre-activations
        # in ANY fp32 evaluation -- the PyTorch yardstick included, at other pixels -- and each flip moves a channel's
        # gradient by up to a few per cent, so a multiple of the yardstick's own (random) flip error is not a bound there.
        # tests/test_gpu_conv.py holds those stacks to 2e-3 against the oracle evaluated on shared masks; here 5e-2.
        floor_k = max(grad_floor, 5e-2) if k.startswith(("spm.", "dec.")) else grad_floor
        if not e_ours < max(floor_k, grad_factor * e_t32):
            fails.append((k, e_ours, e_t32))
    worst.sort(reverse=True)
    rep["worst_grads"] = [dict(param=k, ours=a, torch_same_precision=b) for a, b, k in worst[:8]]
    import statistics
    rep["grad_error_median"] = dict(ours=statistics.median(a for a, _, _ in worst), torch_same_precision=statistics.median(b for _, b, _ in worst))
    rep["n_grads"] = len(worst)
    fs["report"][mode] = rep
    print(f"[{fs['name']} {mode}] activations: " + ", ".join(f"{k} {v['ours']:.1e}" for k, v in rep["activations"].items()))
    print(f"[{fs['name']} {mode}] mask: {rep['mask']}")
    print(f"[{fs['name']} {mode}] gradient error medians (ours / torch at the same precision): {rep['grad_error_median']}")
    print(f"[{fs['name']} {mode}] worst gradients (ours vs torch at the same precision, both against fp64): "
          + "; ".join(f"{k} {a:.1e}/{b:.1e}" for a, b, k in worst[:4]))
    rep["failures"] = [str(f) for f in fails]
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"parity_fullsize_{fs['name']}.json"), "w") as f:
            json.dump(fs["report"], f, indent=1)
    assert not fails, (mode, len(fails), fails[:12])
    return rep


def test_fullsize_fp32_mode(fullsize):
    # fp32 mode: activations 1e-4; gradients no further from fp64 than 4x what plain-PyTorch fp32 is (floor 1e-4)
    rep = _compare(fullsize, "fp32", 1e-4, 1e-4, 4.0)
    assert rep["mask"]["flips"] <= 5, rep["mask"]


def test_fullsize_bf16_mode(fullsize):
    # bf16 (performance / benchmarked) mode: north_star's 2e-2 on activations and logits.  Gradients: within 2e-2, or --
    # where bf16 rounding moves plain PyTorch's own gradients (same graph under torch.autocast(bfloat16)) further than
    # that from the fp64 evaluation -- no further than 3x the distance PyTorch bf16 has (both are noise of the same
    # size: measured medians 0.164 vs 0.169 at ViT-L, 0.071 vs 0.080 at ViT-B; a ratio of two noisy numbers).
    _compare(fullsize, "bf16", 2e-2, 2e-2, 3.0)


def test_config3_masktrans_inference(fullsize):
    """BASELINE config[3]: ViT-L/14 + adapters + MaskTransformer(n_cls=8) multi-class INFERENCE on an EndoVis2018-shaped
    synthetic frame (1280 x 1024 RGB resized to 588 x 588 bilinear, tools/dataset.py:106-108), against the oracle
    (adapter encoder + mask transformer + resize + arg-max) evaluated on the GPU in fp64.  fp32 mode: logits 1e-4, the
    8-class arg-max mask bit-exact wherever the top-2 margin exceeds the tolerance granted to the logits (flips reported);
    bf16 mode: 2e-2."""
    if fullsize["name"] != "vit_large":
        pytest.skip("config[3] is quoted on ViT-L/14")
    import torch.nn.functional as F
    from adaptersis_b200.masktrans import MaskTransformer
    from oracle import masktrans as o_mt
    ts = fullsize["ts"]
    enc = ts.encoder
    C = enc.model.embed_dim
    gen = torch.Generator().manual_seed(77)
    frame = torch.rand(1, 3, 1024, 1280, generator=gen)
    img = F.interpolate(frame, size=(588, 588), mode="bilinear").to(DEV)
    torch.manual_seed(77)
    mt = MaskTransformer(n_cls=8, patch_size=14, d_encoder=C, n_layers=2, n_heads=C // 64, d_model=C, d_ff=4 * C,
                         drop_path_rate=0.0, dropout=0.1).to(DEV).eval()
    with torch.no_grad():
        for name, p in mt.named_parameters():          # away from the (1, 0) / 0.02 initialisation: every class gets signal
            r = torch.randn(p.shape, generator=gen).to(DEV)
            if name == "cls_emb":
                p.copy_(0.5 * r)
            elif name.endswith("bias"):
                p.copy_(0.05 * r)
            elif "norm" in name:
                p.copy_(1.0 + 0.1 * r)
    sds = {"vit": enc.model.state_dict(), "spm": enc.backbone_encoder.state_dict(), "inj": enc.cross_vit.state_dict(),
           "ext": enc.cross_cnn.state_dict()}
    w = {tag: {k: (v.detach().double() if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
         for tag, sd in sds.items()}
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            res = o_enc.adapter_encoder(w["vit"], w["spm"], w["inj"], w["ext"], img.double(), enc.model.num_heads)
            sd64 = {k: v.detach().double() for k, v in mt.state_dict().items()}
            masks64 = o_mt.mask_transformer(sd64, res["x"], (588, 588), 14, 8, C // 64)
            out64, pred64 = o_mt.segment(masks64, (588, 588))
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    top2 = out64.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    scale = float(out64.abs().max())
    report = {}
    for mode, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
        with asis.precision(mode), torch.no_grad():
            x = enc(img)["x"]
            masks = mt(x, (588, 588))
            out = F.interpolate(masks.float(), size=(588, 588), mode="bilinear")
            pred = torch.softmax(out, 1).argmax(1)
        e = relerr(out, out64)
        flips = pred != pred64
        decided = margin > 2 * tol * scale
        report[mode] = dict(logits_relerr=e, classes_present=int(pred64.unique().numel()), pixels=int(flips.numel()), flips=int(flips.sum()),
                            flips_at_determined_pixels=int((flips & decided).sum()),
                            max_margin_at_flips=float(margin[flips].max()) if bool(flips.any()) else 0.0)
        print(f"[config3 {mode}] {report[mode]}")
        assert e < tol, (mode, e)
        assert report[mode]["flips_at_determined_pixels"] == 0, report[mode]
    assert report["fp32"]["flips"] <= 5, report["fp32"]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_config3_masktrans.json"), "w") as f:
            json.dump(report, f, indent=1)


def test_weight_cache_follows_parameter_updates():
    """the bf16 operand cache must refresh after optimizer.step() / load_state_dict(), and writes
    through .data are caught by invalidate_weight_cache()."""
    torch.manual_seed(0)
    lin = torch.nn.Linear(64, 64).to(DEV)
    x = torch.randn(8, 64, device=DEV)
    with asis.precision("bf16"):
        y0 = asis.functional.linear(x, lin.weight, lin.bias, out_dtype=torch.float32)
        opt = torch.optim.SGD(lin.parameters(), lr=0.5)
        y0.sum().backward()
        opt.step()
        y1 = asis.functional.linear(x, lin.weight, lin.bias, out_dtype=torch.float32)
        assert not torch.equal(y0, y1)
        ref = x.bfloat16().float() @ lin.weight.detach().bfloat16().float().t() + lin.bias.detach()
        assert relerr(y1, ref) < 2e-3
        sd = {k: v * 0.5 for k, v in lin.state_dict().items()}
        lin.load_state_dict(sd)
        y2 = asis.functional.linear(x, lin.weight, lin.bias, out_dtype=torch.float32)
        assert relerr(y2, x.bfloat16().float() @ lin.weight.detach().bfloat16().float().t() + lin.bias.detach()) < 2e-3
        lin.weight.data.mul_(2.0)                    # no version bump: the documented blind spot ...
        asis.invalidate_weight_cache()               # ... and its remedy
        y3 = asis.functional.linear(x, lin.weight, lin.bias, out_dtype=torch.float32)
        assert relerr(y3, x.bfloat16().float() @ lin.weight.detach().bfloat16().float().t() + lin.bias.detach()) < 2e-3
    n_before = len(asis.functional._wcache)
    del lin, opt, y0, y1, y2, y3                     # (the outputs' graphs hold the parameter)
    import gc
    gc.collect()
    assert len(asis.functional._wcache) < n_before   # entries die with their parameters


def test_step_frames_equals_step_on_the_host_pipeline_tensors():
    # uint8 HWC frames ingested on the device give the SAME training step as the float batch the reference's
    # dataset builds on the host (tools/dataset.py:111-118): the input tensors are bit-identical (test_gpu_dense.py), so the
    # loss is bit-equal; the parameters after the update agree to fp32 rounding (the loss's F.interpolate backward is an
    # ATen kernel with floating-point atomics: two runs of the SAME feed differ in the last bits too)
    g = torch.Generator().manual_seed(31)
    frames = torch.randint(0, 256, (2, 588, 588, 3), generator=g, dtype=torch.uint8)
    masks = torch.randint(0, 2, (2, 588, 588), generator=g, dtype=torch.uint8)
    img_host = (frames.permute(0, 3, 1, 2) / 255.0).contiguous()
    tgt_host = masks.long()
    outs = []
    for feed in ("float", "u8"):
        torch.manual_seed(5)
        ts = TrainStep(arch="vit_small", adapter_heads=6, device=DEV, precision="fp32")
        if feed == "float":
            loss = ts.step(img_host.pin_memory(), tgt_host.pin_memory())
        else:
            loss = ts.step_frames(frames.pin_memory(), masks.pin_memory())
        outs.append((loss, ts.seg_decoder.final_out.weight.detach().clone(), ts.encoder.cross_cnn.attn.value_proj.weight.detach().clone()))
    assert outs[0][0] == outs[1][0]
    assert relerr(outs[0][1], outs[1][1]) < 1e-5 and relerr(outs[0][2], outs[1][2]) < 1e-5
