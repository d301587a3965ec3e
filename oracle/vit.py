"""Oracle: DinoVisionTransformer front/back ends (test infrastructure only).

  * ``patch_embed``             <- ``PatchEmbed.forward`` (dinov2/layers/patch_embed.py:65-81):
                                   stride-p conv == per-patch linear map
  * ``interpolate_pos_encoding``<- dinov2/models/vision_transformer.py:164-188
  * ``prepare_tokens``          <- ``prepare_tokens_with_masks`` (:190-199), masks=None
  * ``get_intermediate_layers`` <- (:237-247, :263-287), non-chunked
  * ``ARCHS``                   <- ``vit_small/base/large`` factories (:305-345)
"""
import math

import torch
import torch.nn.functional as F

from . import layers

ARCHS = {
    "vit_small": dict(embed_dim=384, depth=12, num_heads=6),
    "vit_base": dict(embed_dim=768, depth=12, num_heads=12),
    "vit_large": dict(embed_dim=1024, depth=24, num_heads=16),
}


def patch_embed(sd, prefix, img, patch):
    B, Cin, H, W = img.shape
    assert H % patch == 0, f"Input image height {H} is not a multiple of patch height {patch}"
    assert W % patch == 0, f"Input image width {W} is not a multiple of patch width: {patch}"
    w = sd[prefix + "proj.weight"]                      # [C, Cin, p, p]
    C = w.shape[0]
    gh, gw = H // patch, W // patch
    cols = img.view(B, Cin, gh, patch, gw, patch).permute(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, Cin * patch * patch)
    return F.linear(cols, w.view(C, -1), sd[prefix + "proj.bias"])


def interpolate_pos_encoding(pos_embed, n_patch_tokens, w, h, patch):
    N = pos_embed.shape[1] - 1
    if n_patch_tokens == N and w == h:
        return pos_embed
    pe = pos_embed if pos_embed.dtype == torch.float64 else pos_embed.float()   # (:170; fp64 kept for the fp64 oracle runs)
    cls_pe, patch_pe = pe[:, :1], pe[:, 1:]
    dim = pe.shape[-1]
    side = int(math.sqrt(N))
    w0, h0 = w // patch + 0.1, h // patch + 0.1
    grid = patch_pe.reshape(1, side, side, dim).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    assert int(w0) == grid.shape[-2] and int(h0) == grid.shape[-1]
    return torch.cat([cls_pe, grid.permute(0, 2, 3, 1).reshape(1, -1, dim)], 1).to(pos_embed.dtype)


def prepare_tokens(sd, img, patch):
    B, _, w, h = img.shape
    x = patch_embed(sd, "patch_embed.", img, patch)
    x = torch.cat([sd["cls_token"].expand(B, -1, -1), x], 1)
    return x + interpolate_pos_encoding(sd["pos_embed"], x.shape[1] - 1, w, h, patch)


def depth_of(sd):
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def get_intermediate_layers(sd, img, n, num_heads, patch, return_class_token=False, norm=True, reshape=False):
    depth = depth_of(sd)
    take = range(depth - n, depth) if isinstance(n, int) else n
    x = prepare_tokens(sd, img, patch)
    outs = []
    for i in range(depth):
        x = layers.block(sd, f"blocks.{i}.", x, num_heads)
        if i in take:
            outs.append(x)
    if norm:
        outs = [layers.layer_norm(sd, "norm.", o) for o in outs]
    cls = [o[:, 0] for o in outs]
    outs = [o[:, 1:] for o in outs]
    if reshape:
        B, _, w, h = img.shape
        outs = [o.reshape(B, w // patch, h // patch, -1).permute(0, 3, 1, 2).contiguous() for o in outs]
    if return_class_token:
        return tuple(zip(outs, cls))
    return tuple(outs)
