"""DINOv2 building blocks with the reference's class names, constructor arguments, forward
signatures and state_dict keys (dinov2/layers/{block,attention,mlp,layer_scale,patch_embed}.py),
executing on libasis_b200 kernels.  Dropout / stochastic depth are identities for the eval-built
backbone the reference uses (drop_path=0, p=0) and are rejected otherwise."""
from typing import Callable, Optional, Tuple, Union

import torch
from torch import Tensor, nn

from . import functional as Fn


def _no_dropout(p, what):
    if p and p > 0.0:
        raise NotImplementedError(f"{what} > 0 is not on the AdapterSIS hot path (reference builds with 0)")


class LayerScale(nn.Module):
    """dinov2/layers/layer_scale.py:15-27 -- x * gamma."""

    def __init__(self, dim: int, init_values: Union[float, Tensor] = 1e-5, inplace: bool = False) -> None:
        super().__init__()
        self.inplace = inplace
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x: Tensor) -> Tensor:
        # stand-alone use only; inside Block the scale is a GEMM epilogue
        return x.mul_(self.gamma) if self.inplace else x * self.gamma


class Mlp(nn.Module):
    """dinov2/layers/mlp.py:16-40."""

    def __init__(self, in_features: int, hidden_features: Optional[int] = None, out_features: Optional[int] = None,
                 act_layer: Callable[..., nn.Module] = nn.GELU, drop: float = 0.0, bias: bool = True) -> None:
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("only the exact-erf nn.GELU activation is implemented")
        _no_dropout(drop, "Mlp drop")
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop = nn.Dropout(drop)

    def forward(self, x: Tensor) -> Tensor:
        return Fn.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


class Attention(nn.Module):
    """dinov2/layers/attention.py:36-69."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False, proj_bias: bool = True,
                 attn_drop: float = 0.0, proj_drop: float = 0.0) -> None:
        super().__init__()
        _no_dropout(attn_drop, "attn_drop")
        _no_dropout(proj_drop, "proj_drop")
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x: Tensor) -> Tensor:
        qkv = Fn.linear(x, self.qkv.weight, self.qkv.bias)
        o = Fn.attention(qkv, self.num_heads)
        return Fn.linear(o, self.proj.weight, self.proj.bias)


class MemEffAttention(Attention):
    """dinov2/layers/attention.py:72-89.  The memory-efficient (flash-style) kernel is ours."""

    def forward(self, x: Tensor, attn_bias=None) -> Tensor:
        if attn_bias is not None:
            raise AssertionError("xFormers is required for using nested tensors")
        return super().forward(x)


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        _no_dropout(drop_prob, "drop_path")
        self.drop_prob = drop_prob

    def forward(self, x):
        return x


class Block(nn.Module):
    """dinov2/layers/block.py:43-114."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = False,
                 proj_bias: bool = True, ffn_bias: bool = True, drop: float = 0.0, attn_drop: float = 0.0,
                 init_values=None, drop_path: float = 0.0, act_layer: Callable[..., nn.Module] = nn.GELU,
                 norm_layer: Callable[..., nn.Module] = nn.LayerNorm, attn_class: Callable[..., nn.Module] = Attention,
                 ffn_layer: Callable[..., nn.Module] = Mlp) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = attn_class(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop,
                               proj_drop=drop)
        self.ls1 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = ffn_layer(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop,
                             bias=ffn_bias)
        self.ls2 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.sample_drop_ratio = drop_path
        if not isinstance(self.norm1, nn.LayerNorm) or not isinstance(self.mlp, Mlp):
            raise NotImplementedError("Block is implemented for LayerNorm + Mlp (ffn_layer='mlp')")

    def forward(self, x: Tensor) -> Tensor:
        g1 = self.ls1.gamma if isinstance(self.ls1, LayerScale) else None
        g2 = self.ls2.gamma if isinstance(self.ls2, LayerScale) else None
        a, m = self.attn, self.mlp
        return Fn.BlockFunction.apply(
            x, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, g1,
            self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, g2,
            a.num_heads, self.norm1.eps, Fn.get_precision(), torch.is_grad_enabled())


class NestedTensorBlock(Block):
    """dinov2/layers/block.py:211-260: list inputs need xformers' BlockDiagonalMask; the AdapterSIS
    path only ever passes tensors."""

    def forward(self, x_or_x_list):
        if isinstance(x_or_x_list, Tensor):
            return super().forward(x_or_x_list)
        raise AssertionError("xFormers is required for using nested tensors")


def make_2tuple(x):
    if isinstance(x, tuple):
        assert len(x) == 2
        return x
    assert isinstance(x, int)
    return (x, x)


class PatchEmbed(nn.Module):
    """dinov2/layers/patch_embed.py:25-81: (B,C,H,W) -> (B,N,D)."""

    def __init__(self, img_size: Union[int, Tuple[int, int]] = 224, patch_size: Union[int, Tuple[int, int]] = 16,
                 in_chans: int = 3, embed_dim: int = 768, norm_layer: Optional[Callable] = None,
                 flatten_embedding: bool = True) -> None:
        super().__init__()
        image_HW = make_2tuple(img_size)
        patch_HW = make_2tuple(patch_size)
        if patch_HW[0] != patch_HW[1]:
            raise NotImplementedError("square patches only")
        self.img_size = image_HW
        self.patch_size = patch_HW
        self.patches_resolution = (image_HW[0] // patch_HW[0], image_HW[1] // patch_HW[1])
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.flatten_embedding = flatten_embedding
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_HW, stride=patch_HW)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x: Tensor) -> Tensor:
        _, _, H, W = x.shape
        patch_H, patch_W = self.patch_size
        assert H % patch_H == 0, f"Input image height {H} is not a multiple of patch height {patch_H}"
        assert W % patch_W == 0, f"Input image width {W} is not a multiple of patch width: {patch_W}"
        x = Fn.PatchEmbedFunction.apply(x, self.proj.weight, self.proj.bias, patch_H, Fn.get_precision())
        if isinstance(self.norm, nn.LayerNorm):
            x = Fn.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps, torch.float32)
        if not self.flatten_embedding:
            x = x.reshape(-1, H // patch_H, W // patch_W, self.embed_dim)
        return x
