"""Mask-transformer decoder (Segmenter-style) of the reference's multi-class evaluation scripts -- BASELINE.json
config[3]: ``MaskTransformer`` (eval/eval_dinov2_masktrans.py:400-465) over ``Block`` / ``Attention`` / ``FeedForward``
(backbones/masktrans_block.py:11-95) -- with the reference's class names, constructor arguments, forward signatures and
state_dict keys, executing on libasis_b200 kernels:

  proj_dec                        one GEMM (bias epilogue)
  2 x Block                       the single-node transformer block of the backbone (functional.BlockFunction: LayerNorm,
                                  fused QKV GEMM, flash attention over 1764 + n_cls tokens, projection + residual epilogue,
                                  fc1 + exact GELU epilogue, fc2 + residual epilogue) with LayerNorm eps 1e-5 and no LayerScale
  decoder_norm                    LayerNorm kernel
  patches @ proj_patch,           GEMMs with the parameter read MN-major (it is stored [d_in, d_out]: no transpose)
  cls_seg_feat @ proj_classes
  L2 normalisation, masks = patches @ cls^T (an n_cls-wide product), LayerNorm over n_cls, rearrange:
                                  a few MB of elementwise work, left to PyTorch device ops (not on the hot path)

Dropout (p = 0.1 in the reference's constructor call) is the identity in ``eval()``; in training mode it is rejected
rather than silently skipped.  ``return_attention`` (a visualisation path that materialises the T x T matrix) is not
provided."""
import torch
import torch.nn as nn

from . import functional as Fn
from . import kernels as K
from ._lib import MAJOR_K, MAJOR_MN


def _check_dropout(mod):
    if mod.training and mod.p > 0.0:
        raise NotImplementedError("MaskTransformer: dropout > 0 in training mode is not implemented (inference path: call .eval())")


class FeedForward(nn.Module):
    """backbones/masktrans_block.py:11-31."""

    def __init__(self, dim, hidden_dim, dropout, out_dim=None):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden_dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_dim, dim if out_dim is None else out_dim)
        self.drop = nn.Dropout(dropout)

    @property
    def unwrapped(self):
        return self

    def forward(self, x):
        _check_dropout(self.drop)
        return Fn.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


class Attention(nn.Module):
    """backbones/masktrans_block.py:34-72 (returns ``(x, None)``: the attention matrix is never materialised)."""

    def __init__(self, dim, heads, dropout):
        super().__init__()
        self.heads = heads
        self.scale = (dim // heads) ** -0.5
        self.attn = None
        self.qkv = nn.Linear(dim, dim * 3)
        self.attn_drop = nn.Dropout(dropout)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(dropout)

    @property
    def unwrapped(self):
        return self

    def forward(self, x, mask=None):
        _check_dropout(self.attn_drop)
        qkv = Fn.linear(x, self.qkv.weight, self.qkv.bias)
        o = Fn.attention(qkv, self.heads)
        return Fn.linear(o, self.proj.weight, self.proj.bias), None


class Block(nn.Module):
    """backbones/masktrans_block.py:75-95: x + attn(norm1 x); x + mlp(norm2 x) -- one autograd node."""

    def __init__(self, dim, heads, mlp_dim, dropout, drop_path):
        super().__init__()
        if drop_path > 0.0:
            raise NotImplementedError("MaskTransformer Block: drop_path > 0 is not implemented (the reference passes 0.0)")
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn = Attention(dim, heads, dropout)
        self.mlp = FeedForward(dim, mlp_dim, dropout)
        self.drop_path = nn.Identity()

    def forward(self, x, mask=None, return_attention=False):
        if return_attention:
            raise NotImplementedError("return_attention materialises the T x T attention matrix: visualisation only, not provided")
        a, m = self.attn, self.mlp
        _check_dropout(a.attn_drop)
        return Fn.BlockFunction.apply(
            x, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, None,
            self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, None,
            a.heads, self.norm1.eps, Fn.get_precision(), torch.is_grad_enabled())


def init_weights(m):
    """eval/eval_dinov2_masktrans.py:389-396."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class _ParamMatmul(torch.autograd.Function):
    """x [R, K] @ P [K, N] with P a parameter stored [K, N]: the GEMM reads it MN-major, no transposed copy."""

    @staticmethod
    def forward(ctx, x, P, mode):
        comp, cdt = Fn._cfg(mode)
        shp = x.shape
        x2 = K.cast(x.reshape(-1, shp[-1]).contiguous(), cdt)
        Pc = Fn._operand(P, cdt)
        y, _ = K.gemm(comp, x2, MAJOR_K, Pc, MAJOR_MN, x2.shape[0], P.shape[1], P.shape[0], torch.float32)
        ctx.save_for_backward(x2, P)
        ctx.meta = (shp, mode, x.dtype)
        return y.view(*shp[:-1], P.shape[1])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x2, P = ctx.saved_tensors
        shp, mode, xdt = ctx.meta
        comp, cdt = Fn._cfg(mode)
        dy2 = K.cast(dy.reshape(-1, P.shape[1]).contiguous(), cdt)
        dx = dP = None
        if ctx.needs_input_grad[0]:          # dx = dy P^T: P [K, N] is the K-major "weight" of this product
            dx, _ = K.gemm(comp, dy2, MAJOR_K, Fn._operand(P, cdt), MAJOR_K, dy2.shape[0], P.shape[0], P.shape[1], torch.float32)
            dx = dx.view(shp).to(xdt)
        if ctx.needs_input_grad[1]:          # dP [K, N] = x^T dy
            dP, _ = K.gemm(comp, x2, MAJOR_MN, dy2, MAJOR_MN, P.shape[0], P.shape[1], x2.shape[0], torch.float32)
            dP = dP.to(P.dtype)
        return dx, dP, None


class MaskTransformer(nn.Module):
    """eval/eval_dinov2_masktrans.py:400-465.  forward(x [B, N, d_encoder], im_size) -> masks [B, n_cls, H/p, W/p]."""

    def __init__(self, n_cls, patch_size, d_encoder, n_layers, n_heads, d_model, d_ff, drop_path_rate, dropout):
        super().__init__()
        self.d_encoder = d_encoder
        self.patch_size = patch_size
        self.n_layers = n_layers
        self.n_cls = n_cls
        self.d_model = d_model
        self.d_ff = d_ff
        self.scale = d_model ** -0.5
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, n_layers)]
        self.blocks = nn.ModuleList([Block(d_model, n_heads, d_ff, dropout, dpr[i]) for i in range(n_layers)])
        self.cls_emb = nn.Parameter(torch.randn(1, n_cls, d_model))
        self.proj_dec = nn.Linear(d_encoder, d_model)
        self.proj_patch = nn.Parameter(self.scale * torch.randn(d_model, d_model))
        self.proj_classes = nn.Parameter(self.scale * torch.randn(d_model, d_model))
        self.decoder_norm = nn.LayerNorm(d_model)
        self.mask_norm = nn.LayerNorm(n_cls)
        self.apply(init_weights)
        nn.init.trunc_normal_(self.cls_emb, std=0.02)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"cls_emb"}

    def forward(self, x, im_size):
        H, W = im_size
        GS = H // self.patch_size
        mode = Fn.get_precision()
        x = Fn.linear(x, self.proj_dec.weight, self.proj_dec.bias, out_dtype=torch.float32)
        cls_emb = self.cls_emb.expand(x.size(0), -1, -1)
        x = torch.cat((x, cls_emb.to(x.dtype)), 1)
        for blk in self.blocks:
            x = blk(x)
        dn = self.decoder_norm
        x = Fn.layer_norm(x, dn.weight, dn.bias, dn.eps, out_dtype=torch.float32)
        patches, cls_seg_feat = x[:, : -self.n_cls], x[:, -self.n_cls:]
        patches = _ParamMatmul.apply(patches, self.proj_patch, mode)
        cls_seg_feat = _ParamMatmul.apply(cls_seg_feat, self.proj_classes, mode)
        # the n_cls-wide tail: a few MB of elementwise work and one [N, d] x [d, n_cls] product, fp32 device ops
        patches = patches / patches.norm(dim=-1, keepdim=True)
        cls_seg_feat = cls_seg_feat / cls_seg_feat.norm(dim=-1, keepdim=True)
        masks = patches @ cls_seg_feat.transpose(1, 2)
        masks = self.mask_norm(masks)
        B, N, _ = masks.shape
        return masks.view(B, int(GS), N // int(GS), self.n_cls).permute(0, 3, 1, 2)

    def get_attention_map(self, x, layer_id):
        raise NotImplementedError("get_attention_map materialises the T x T attention matrix: visualisation only, not provided")
