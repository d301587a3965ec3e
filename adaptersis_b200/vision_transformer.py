"""DinoVisionTransformer with the reference's constructor, attributes, methods and state_dict keys
(dinov2/models/vision_transformer.py:44-357), running on libasis_b200 kernels.

Same object model as the reference so that public DINOv2 checkpoints load key for key and
``train.py``-style code can keep indexing ``model.blocks[i]`` / calling ``model.patch_embed``; what
runs underneath is different:

* ``patch_embed`` is a patch gather (``asis_patchify``) plus one tcgen05 GEMM instead of a strided
  convolution; its backward is the same GEMM with the other operand majors;
* a block is ONE autograd node (``functional.BlockFunction``): LayerNorm, fused-QKV GEMM, flash-style
  attention over the 1765 tokens, projection + LayerScale + residual epilogue, LayerNorm, fc1 + GELU
  epilogue, fc2 + LayerScale + residual epilogue -- seven kernels forward; the residual stream,
  LayerNorm statistics and LayerScale stay fp32 in both precision modes;
* the bicubic interpolation of the position embedding (37 x 37 -> 42 x 42 at 588^2) is static per
  resolution and cached per parameter version instead of being recomputed in every forward;
* forward-only calls (``get_intermediate_layers`` under ``inference_mode`` -- the taps pass of the
  training step) skip every tensor that exists only for backward.
"""
import math
from functools import partial
from typing import Sequence, Tuple, Union

import torch
import torch.nn as nn
from torch.nn.init import trunc_normal_

from . import functional as Fn
from .layers import MemEffAttention, Mlp, NestedTensorBlock as Block, PatchEmbed


class BlockChunk(nn.ModuleList):
    """One FSDP wrapping unit of the reference (vision_transformer.py:37-41): runs its members in order."""

    def forward(self, x):
        for b in self:
            x = b(x)
        return x


class DinoVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=True, ffn_bias=True, proj_bias=True, drop_path_rate=0.0, drop_path_uniform=False,
                 init_values=None, embed_layer=PatchEmbed, act_layer=nn.GELU, block_fn=Block, ffn_layer="mlp",
                 block_chunks=1):
        super().__init__()
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 1
        self.n_blocks = depth
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + self.num_tokens, embed_dim))
        if drop_path_rate != 0.0:
            raise NotImplementedError("drop_path_rate > 0 is not on the AdapterSIS hot path")
        if ffn_layer != "mlp":
            raise NotImplementedError("ffn_layer must be 'mlp' (configs/eval/vitl14_pretrain.yaml)")
        blocks = [block_fn(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                           proj_bias=proj_bias, ffn_bias=ffn_bias, drop_path=0.0, norm_layer=norm_layer,
                           act_layer=act_layer, ffn_layer=Mlp, init_values=init_values) for _ in range(depth)]
        if block_chunks > 0:
            # FSDP-style chunk layout (:141-147): chunk c holds blocks [c*size, (c+1)*size) at their GLOBAL
            # positions (identity placeholders in front), so checkpoints saved with ``blocks.<c>.<i>.`` keys load
            # key for key.  train.py builds with block_chunks=0 and indexes model.blocks[i] directly.
            self.chunked_blocks = True
            size = depth // block_chunks
            self.blocks = nn.ModuleList([BlockChunk([nn.Identity()] * i + blocks[i:i + size])
                                         for i in range(0, depth, size)])
        else:
            self.chunked_blocks = False
            self.blocks = nn.ModuleList(blocks)
        self.norm = norm_layer(embed_dim)
        self.head = nn.Identity()
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self._pos_cache = {}
        self.init_weights()

    def init_weights(self):
        trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def interpolate_pos_encoding(self, x, w, h):
        """Bicubic resize of the patch position grid (:164-188).  Constant per resolution, so the
        result is cached while pos_embed is frozen (host plumbing, not a hot kernel)."""
        previous_dtype = x.dtype
        npatch = x.shape[1] - 1
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        key = (w, h, self.pos_embed._version, self.pos_embed.device)
        if not self.pos_embed.requires_grad or not torch.is_grad_enabled():
            hit = self._pos_cache.get(key)
            if hit is not None:
                return hit.to(previous_dtype)
        pos_embed = self.pos_embed.float()
        class_pos_embed = pos_embed[:, 0]
        patch_pos_embed = pos_embed[:, 1:]
        dim = x.shape[-1]
        w0 = w // self.patch_size + 0.1
        h0 = h // self.patch_size + 0.1
        side = int(math.sqrt(N))
        patch_pos_embed = nn.functional.interpolate(
            patch_pos_embed.reshape(1, side, side, dim).permute(0, 3, 1, 2),
            scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
        assert int(w0) == patch_pos_embed.shape[-2] and int(h0) == patch_pos_embed.shape[-1]
        patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).view(1, -1, dim)
        out = torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)
        if not out.requires_grad:
            self._pos_cache = {key: out}
        return out.to(previous_dtype)

    def prepare_tokens_with_masks(self, x, masks=None):
        B, nc, w, h = x.shape
        x = self.patch_embed(x)
        if masks is not None:
            x = torch.where(masks.unsqueeze(-1), self.mask_token.to(x.dtype).unsqueeze(0), x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1)
        x = x + self.interpolate_pos_encoding(x, w, h)
        return x

    def _norm(self, x):
        return Fn.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps, torch.float32)

    def forward_features(self, x, masks=None):
        if isinstance(x, list):
            return [self.forward_features(xi, mi) for xi, mi in zip(x, masks)]
        x = self.prepare_tokens_with_masks(x, masks)
        for blk in self.blocks:          # a Block, or a BlockChunk running its members (identities in front)
            x = blk(x)
        x_norm = self._norm(x)
        return {"x_norm_clstoken": x_norm[:, 0], "x_norm_patchtokens": x_norm[:, 1:], "x_prenorm": x, "masks": masks}

    def _real_blocks(self):
        """The transformer blocks in execution order, whichever layout ``self.blocks`` has."""
        if not self.chunked_blocks:
            return list(self.blocks)
        return [b for chunk in self.blocks for b in chunk if not isinstance(b, nn.Identity)]

    def _get_intermediate_layers_not_chunked(self, x, n=1):
        # serves both layouts (:237-261): outputs are indexed by global block position
        x = self.prepare_tokens_with_masks(x)
        blocks = self._real_blocks()
        output, total_block_len = [], len(blocks)
        blocks_to_take = range(total_block_len - n, total_block_len) if isinstance(n, int) else n
        for i, blk in enumerate(blocks):
            x = blk(x)
            if i in blocks_to_take:
                output.append(x)
        assert len(output) == len(blocks_to_take), f"only {len(output)} / {len(blocks_to_take)} blocks found"
        return output

    _get_intermediate_layers_chunked = _get_intermediate_layers_not_chunked

    def get_intermediate_layers(self, x: torch.Tensor, n: Union[int, Sequence] = 1, reshape: bool = False,
                                return_class_token: bool = False, norm=True) -> Tuple[Union[torch.Tensor, Tuple[torch.Tensor]]]:
        outputs = self._get_intermediate_layers_not_chunked(x, n)
        if norm:
            outputs = [self._norm(out) for out in outputs]
        class_tokens = [out[:, 0] for out in outputs]
        outputs = [out[:, 1:] for out in outputs]
        if reshape:
            B, _, w, h = x.shape
            outputs = [out.reshape(B, w // self.patch_size, h // self.patch_size, -1).permute(0, 3, 1, 2).contiguous()
                       for out in outputs]
        if return_class_token:
            return tuple(zip(outputs, class_tokens))
        return tuple(outputs)

    def forward(self, *args, is_training=False, **kwargs):
        ret = self.forward_features(*args, **kwargs)
        if is_training:
            return ret
        return self.head(ret["x_norm_clstoken"])


def _vit(patch_size, embed_dim, depth, num_heads, **kwargs):
    return DinoVisionTransformer(patch_size=patch_size, embed_dim=embed_dim, depth=depth, num_heads=num_heads,
                                 mlp_ratio=4, block_fn=partial(Block, attn_class=MemEffAttention), **kwargs)


def vit_small(patch_size=16, **kwargs):
    return _vit(patch_size, 384, 12, 6, **kwargs)


def vit_base(patch_size=16, **kwargs):
    return _vit(patch_size, 768, 12, 12, **kwargs)


def vit_large(patch_size=16, **kwargs):
    return _vit(patch_size, 1024, 24, 16, **kwargs)


def vit_giant2(patch_size=16, **kwargs):
    return _vit(patch_size, 1536, 40, 24, **kwargs)


def build_model_for_eval(arch="vit_large", img_size=518, patch_size=14, init_values=1e-5, **kw):
    """What dinov2/eval/setup.py:62-67 + dinov2/models/__init__.py:14-40 produce for
    configs/eval/vit{s,b,l}14_pretrain.yaml: an eval-mode teacher with LayerScale 1e-5, mlp FFN,
    block_chunks 0."""
    fn = {"vit_small": vit_small, "vit_base": vit_base, "vit_large": vit_large, "vit_giant2": vit_giant2}[arch]
    model = fn(patch_size=patch_size, img_size=img_size, init_values=init_values, ffn_layer="mlp", block_chunks=0,
               qkv_bias=True, proj_bias=True, ffn_bias=True, **kw)
    model.eval()
    return model


class ModelWithIntermediateLayers(nn.Module):
    """dinov2/eval/utils.py:30-44."""

    def __init__(self, feature_model, n_last_blocks, autocast_ctx=None):
        super().__init__()
        self.feature_model = feature_model
        self.feature_model.eval()
        self.n_last_blocks = n_last_blocks
        self.autocast_ctx = autocast_ctx

    def forward(self, images):
        with torch.no_grad():
            return self.feature_model.get_intermediate_layers(images, self.n_last_blocks, return_class_token=True)
