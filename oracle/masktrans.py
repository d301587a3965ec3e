"""Oracle: the mask-transformer decoder of the reference's multi-class evaluation (test infrastructure only).

  * ``mask_transformer`` <- ``MaskTransformer.forward`` (eval/eval_dinov2_masktrans.py:443-465) over
                            ``Block.forward`` (backbones/masktrans_block.py:84-95), eval mode (dropout = identity),
                            ``nn.LayerNorm`` default eps 1e-5
  * ``segment``          <- the inference tail of ``validate_network`` (:361-366): bilinear resize to the image size,
                            softmax, arg-max
"""
import torch
import torch.nn.functional as F

from . import layers

LN_EPS = 1e-5


def _block(sd, prefix, x, heads):
    x = x + layers.attention(sd, prefix + "attn.", layers.layer_norm(sd, prefix + "norm1.", x, LN_EPS), heads)
    return x + layers.mlp(sd, prefix + "mlp.", layers.layer_norm(sd, prefix + "norm2.", x, LN_EPS))


def mask_transformer(sd, x, im_size, patch_size, n_cls, n_heads):
    H, W = im_size
    GS = H // patch_size
    x = F.linear(x, sd["proj_dec.weight"], sd["proj_dec.bias"])
    x = torch.cat((x, sd["cls_emb"].expand(x.shape[0], -1, -1)), 1)
    n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    for i in range(n_layers):
        x = _block(sd, f"blocks.{i}.", x, n_heads)
    x = layers.layer_norm(sd, "decoder_norm.", x, LN_EPS)
    patches, cls = x[:, :-n_cls], x[:, -n_cls:]
    patches = patches @ sd["proj_patch"]
    cls = cls @ sd["proj_classes"]
    patches = patches / patches.norm(dim=-1, keepdim=True)
    cls = cls / cls.norm(dim=-1, keepdim=True)
    masks = layers.layer_norm(sd, "mask_norm.", patches @ cls.transpose(1, 2), LN_EPS)
    B, N, _ = masks.shape
    return masks.view(B, GS, N // GS, n_cls).permute(0, 3, 1, 2)


def segment(masks, im_size):
    """logits at the image size, class probabilities' arg-max (eval_dinov2_masktrans.py:362-366)."""
    out = F.interpolate(masks, size=im_size, mode="bilinear")
    return out, torch.softmax(out, dim=1).argmax(1)
