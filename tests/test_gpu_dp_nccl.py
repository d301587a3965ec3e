"""SURVEY.md section 8(e) parity check, on NCCL: the gradients a 2-GPU data-parallel step produces (per-rank
shards, bucketed all-reduce AVG -- after backward, the default, or overlapped with it --, SyncBN statistics exchanged
in the spatial prior module) equal the gradients ONE GPU produces on the same global batch.

The decoder's plain BatchNorm2d uses per-GPU batch statistics under DP by construction (as in the reference:
backbones/decoders.py uses nn.BatchNorm2d under DDP), which is a semantic difference of DP itself, not of the
gradient exchange; the decoder is therefore put in eval() for this comparison so that the only cross-sample
couplings are the SyncBN layers, which DP must reproduce exactly.

Needs >= 2 GPUs: skipped on the single-GPU test box; run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_dp_nccl.py` (result recorded in profiles/)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, 588, 588, generator=g), torch.randint(0, 2, (n, 588, 588), generator=g)


def _make(precision, dev):
    from adaptersis_b200.trainer import TrainStep
    torch.manual_seed(7)
    ts = TrainStep(arch="vit_small", adapter_heads=6, device=dev, precision=precision, bucket_bytes=8 << 20)
    with torch.no_grad():                    # make the injector and LayerScale carry signal
        ts.encoder.cross_vit.gamma.fill_(0.5)
        for blk in ts.encoder.model.blocks:
            blk.ls1.gamma.fill_(0.5)
            blk.ls2.gamma.fill_(0.5)
    ts.seg_decoder.eval()
    return ts


def _grads(ts, img, tgt):
    for p in ts.parameters():
        p.grad = None
    internals = {}
    loss = ts.forward_loss(img, tgt, internals)
    # the dice loss of a freshly initialised 2-class head is almost flat (gradients at fp32-noise level), so a second
    # term drives the encoder graph directly, as in the golden fixtures; it is a per-image mean, like the loss
    feat = internals["feat"]
    g = torch.Generator().manual_seed(5)
    probe = torch.randn(feat.shape[1:], generator=g).to(feat.device)
    loss = loss + (feat * probe).sum() / (feat.shape[0] * feat[0].numel() ** 0.5)
    loss.backward()
    ts.reducer.finish()
    torch.cuda.synchronize()
    return {k: p.grad.detach().float().cpu() for k, p in ts.named_parameters() if p.grad is not None}, float(loss)


def _worker(rank, world, port, precision, per_rank, overlap, q):
    os.environ["ASIS_DP_OVERLAP"] = overlap
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ts = _make(precision, dev)
        assert ts.reducer.world == world and len(ts.reducer.buckets) > 2 and ts.reducer.overlap == (overlap == "1")
        img, tgt = _batch(world * per_rank, 99)
        sl = slice(rank * per_rank, (rank + 1) * per_rank)
        g, loss = _grads(ts, img[sl].to(dev), tgt[sl].to(dev))
        q.put((rank, g if rank == 0 else None, loss))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc(), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision,tol,overlap", [("fp32", 1e-3, "0"), ("bf16", 2e-2, "0"), ("fp32", 1e-3, "1")])
def test_two_gpu_gradients_equal_one_gpu(precision, tol, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world, per_rank = 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, precision, per_rank, overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(60)
    for r, g, loss in res:
        assert loss is not None, g
    g2 = res[0][1]
    loss2 = sum(r[2] for r in res) / world
    # one GPU, the same 4 images
    dev = torch.device("cuda", 0)
    ts = _make(precision, dev)
    img, tgt = _batch(world * per_rank, 99)
    g1, loss1 = _grads(ts, img.to(dev), tgt.to(dev))
    assert abs(loss1 - loss2) < (1e-5 if precision == "fp32" else 2e-3), (loss1, loss2)
    assert set(g1) == set(g2)
    # The spatial prior module's gradients sit behind ReLU'(0) discontinuities: the 2-GPU batch statistics (two partial
    # sums merged with Chan's formula) differ from the 1-GPU ones in the last bit, a handful of pre-activations within
    # rounding distance of zero change branch, and each such flip moves a channel's gradient by up to a few per cent on
    # these small maps (tests/test_gpu_conv.py, tools/spm_debug.py).  Those parameters get 3e-2 (fp32) / 2e-1 (bf16);
    # everything else -- backbone, adapters, decoder: the gradient exchange proper -- the tight bound.
    spm_tol = 3e-2 if precision == "fp32" else 2e-1
    worst, worst_spm, errs = (0.0, None), (0.0, None), []
    for k in g1:
        scale = float(g1[k].abs().max())
        if scale < 1e-12:
            continue
        e = float((g1[k] - g2[k]).abs().max()) / scale
        errs.append((e, k))
        if "backbone_encoder" in k:
            worst_spm = max(worst_spm, (e, k))
        else:
            worst = max(worst, (e, k))
    errs.sort(reverse=True)
    print(f"[dp nccl {precision}] largest: " + "; ".join(f"{k} {e:.1e}" for e, k in errs[:6]) + f"; median {errs[len(errs) // 2][0]:.1e}")
    print(f"[dp nccl {precision}] {len(g1)} gradients, worst 2-GPU vs 1-GPU error {worst[0]:.2e} ({worst[1]})")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"dp_nccl_parity_{precision}{'_overlap' if overlap == '1' else ''}.txt"), "w") as f:
            f.write(f"{len(g1)} gradients; worst error {worst[0]:.3e} at {worst[1]} (spatial prior module: {worst_spm[0]:.3e} at "
                    f"{worst_spm[1]}); loss 1-GPU {loss1:.7f} 2-GPU mean {loss2:.7f}\n")
    assert worst[0] < tol, worst
    assert worst_spm[0] < spm_tol, worst_spm
