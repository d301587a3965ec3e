"""Checkpoints.  The reference saves ONLY the decoder (train.py:248-255:
``{"epoch", "state_dict": seg_decoder.state_dict(), "optimizer", "scheduler", "best_acc"}``, with DDP's ``module.``
prefix on the keys) although its step also runs the spatial prior module and the two adapter blocks -- a reference
checkpoint cannot restore the network it trained.  This module writes a COMPLETE checkpoint that
is a superset of the reference's: the same five top-level keys with the same meaning (so the reference's own loader,
``utils.restart_from_checkpoint``, finds what it looks for), plus the four encoder-side state_dicts under ``"encoder"``
-- every key a name the reference's own modules produce -- and reads both formats.

``load_pretrained_weights`` mirrors dinov2/utils/utils.py:17-32 for the public DINOv2 backbone checkpoints."""
import torch

from . import functional as Fn

FORMAT = "adaptersis_b200/1"


def _strip(sd, *prefixes):
    out = {}
    for k, v in sd.items():
        for p in prefixes:
            k = k.replace(p, "")
        out[k] = v
    return out


def checkpoint_dict(train_step, epoch, best_acc=0.0, scheduler=None):
    enc = train_step.encoder
    return {
        # --- the reference's keys (train.py:248-254)
        "epoch": epoch,
        "state_dict": train_step.seg_decoder.state_dict(),
        "optimizer": train_step.optimizer.state_dict(),
        "scheduler": scheduler.state_dict() if scheduler is not None else None,
        "best_acc": best_acc,
        # --- what the reference leaves out: everything else its train() iteration reads
        "format": FORMAT,
        "encoder": {
            "backbone": enc.model.state_dict(),                  # DinoVisionTransformer keys (vision_transformer.py)
            "backbone_encoder": enc.backbone_encoder.state_dict(),   # FeatureEncoder keys incl. SyncBN running statistics
            "cross_vit": enc.cross_vit.state_dict(),             # CAViT keys
            "cross_cnn": enc.cross_cnn.state_dict(),             # CACNN keys
        },
    }


def save_checkpoint(path, train_step, epoch, best_acc=0.0, scheduler=None):
    torch.save(checkpoint_dict(train_step, epoch, best_acc, scheduler), path)


def load_checkpoint(path_or_dict, train_step, scheduler=None, strict=True, load_optimizer=True):
    """Restore a TrainStep from a complete checkpoint, or from a reference checkpoint (decoder only: the encoder side
    keeps its current weights, as after the reference's own restart).  Returns ``(epoch, best_acc, missing)`` where
    ``missing`` lists the parts the file did not carry."""
    ck = torch.load(path_or_dict, map_location="cpu", weights_only=False) if isinstance(path_or_dict, (str, bytes)) or hasattr(path_or_dict, "read") else path_or_dict
    missing = []
    train_step.seg_decoder.load_state_dict(_strip(ck["state_dict"], "module."), strict=strict)
    enc = train_step.encoder
    parts = ck.get("encoder") or {}
    for name, mod in (("backbone", enc.model), ("backbone_encoder", enc.backbone_encoder), ("cross_vit", enc.cross_vit),
                      ("cross_cnn", enc.cross_cnn)):
        if name in parts:
            mod.load_state_dict(_strip(parts[name], "module."), strict=strict)
        else:
            missing.append(name)
    if load_optimizer and ck.get("optimizer") is not None and not missing:
        train_step.optimizer.load_state_dict(ck["optimizer"])      # (a decoder-only file has another parameter list)
    elif load_optimizer:
        missing.append("optimizer")
    if scheduler is not None and ck.get("scheduler") is not None:
        scheduler.load_state_dict(ck["scheduler"])
    # parameters were written through .data / copy_: bf16 operand copies and a captured step graph are stale
    Fn.invalidate_weight_cache()
    if getattr(train_step, "_graph", None) is not None:
        train_step._replayed = True
    return ck.get("epoch", 0), ck.get("best_acc", 0.0), missing


def load_pretrained_weights(model, pretrained_weights, checkpoint_key=None):
    """dinov2/utils/utils.py:17-32: a DINOv2 backbone checkpoint (file path or an already loaded dict) into
    ``DinoVisionTransformer``: optional ``checkpoint_key`` level, ``module.`` / ``backbone.`` prefixes removed,
    ``strict=False``; returns torch's ``_IncompatibleKeys`` message as the reference logs it."""
    sd = torch.load(pretrained_weights, map_location="cpu") if isinstance(pretrained_weights, str) else pretrained_weights
    if checkpoint_key is not None and checkpoint_key in sd:
        sd = sd[checkpoint_key]
    msg = model.load_state_dict(_strip(sd, "module.", "backbone."), strict=False)
    Fn.invalidate_weight_cache()
    return msg
