"""the complete checkpoint is a superset of the reference's (train.py:248-255), round-trips every
parameter / buffer / optimizer slot, and a reference-format file (decoder only, DDP-prefixed keys) still loads."""
import io

import torch


def _tiny(seed):
    from adaptersis_b200.trainer import TrainStep
    torch.manual_seed(seed)
    return TrainStep(arch="vit_small", adapter_heads=6, device="cpu", precision="fp32", dec_features=[384, 16, 8, 8, 4])


def test_complete_checkpoint_round_trip_and_reference_keys():
    from adaptersis_b200 import checkpoint as ck
    a, b = _tiny(1), _tiny(2)
    # give the optimizer state something to carry
    for p in a.optimizer.param_groups[0]["params"][:3]:
        a.optimizer.state[p]["momentum_buffer"] = torch.randn_like(p)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(a.optimizer, 10, eta_min=0)
    d = ck.checkpoint_dict(a, epoch=3, best_acc=0.5, scheduler=sched)
    assert {"epoch", "state_dict", "optimizer", "scheduler", "best_acc"} <= set(d)          # the reference's five keys
    assert d["state_dict"].keys() == a.seg_decoder.state_dict().keys()
    buf = io.BytesIO()
    torch.save(d, buf)
    buf.seek(0)
    epoch, best, missing = ck.load_checkpoint(buf, b)
    assert (epoch, best, missing) == (3, 0.5, [])
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    pa, pb = a.optimizer.param_groups[0]["params"], b.optimizer.param_groups[0]["params"]
    for x, y in zip(pa[:3], pb[:3]):
        assert torch.equal(a.optimizer.state[x]["momentum_buffer"], b.optimizer.state[y]["momentum_buffer"])


def test_reference_format_checkpoint_loads():
    from adaptersis_b200 import checkpoint as ck
    a, b = _tiny(3), _tiny(4)
    ref = {"epoch": 7, "state_dict": {"module." + k: v for k, v in a.seg_decoder.state_dict().items()},      # DDP-wrapped decoder
           "optimizer": None, "scheduler": None, "best_acc": 0.25}
    before = {k: v.clone() for k, v in b.encoder.state_dict().items()}
    epoch, best, missing = ck.load_checkpoint(ref, b)
    assert epoch == 7 and best == 0.25
    assert set(missing) == {"backbone", "backbone_encoder", "cross_vit", "cross_cnn", "optimizer"}
    for (ka, va), (kb, vb) in zip(a.seg_decoder.state_dict().items(), b.seg_decoder.state_dict().items()):
        assert torch.equal(va, vb), ka
    for k, v in b.encoder.state_dict().items():            # the encoder side is untouched, as after the reference's restart
        assert torch.equal(v, before[k]), k


def test_load_pretrained_weights_strips_prefixes():
    from adaptersis_b200 import checkpoint as ck
    import adaptersis_b200 as asis
    torch.manual_seed(5)
    src = asis.vit_small(patch_size=14, img_size=518, init_values=1e-5, block_chunks=0)
    dst = asis.vit_small(patch_size=14, img_size=518, init_values=1e-5, block_chunks=0)
    wrapped = {"teacher": {"module.backbone." + k: v for k, v in src.state_dict().items()}}
    msg = ck.load_pretrained_weights(dst, wrapped, "teacher")
    assert not msg.missing_keys and not msg.unexpected_keys
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
