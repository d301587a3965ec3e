"""Oracle: injector / extractor adapter blocks (test infrastructure only).

  * ``reference_points`` <- ``get_reference_points`` (backbones/adapter_blocks.py:9-22)
  * ``deform_inputs``    <- ``deform_inputs``        (:24-38)
  * ``dwconv``           <- ``DWConv.forward``       (:62-80)  (pyramid split hard-wired to 18*18, :71)
  * ``conv_ffn``         <- ``ConvFFN.forward``      (:82-100)
  * ``cavit``            <- ``CAViT.forward``  injector (:149-183)
  * ``cacnn``            <- ``CACNN.forward``  extractor (:102-147)
"""
import torch
import torch.nn.functional as F

from . import layers
from .msda import msda_module, msda_core


def reference_points(shapes):
    pts = []
    for H, W in shapes:
        # cell centres (i + 0.5) / size, produced with linspace exactly as the reference does
        ys = torch.linspace(0.5, H - 0.5, H, dtype=torch.float32) / H
        xs = torch.linspace(0.5, W - 0.5, W, dtype=torch.float32) / W
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([gx.reshape(-1), gy.reshape(-1)], -1)[None])
    return torch.cat(pts, 1)[:, :, None]


def deform_inputs(h, w, patch):
    pyr = [(h // 8, w // 8), (h // 16, w // 16), (h // 32, w // 32)]
    grid = [(h // patch, w // patch)]

    def pack(ref_shapes, val_shapes):
        ss = torch.as_tensor(val_shapes, dtype=torch.long)
        lsi = torch.cat([ss.new_zeros(1), ss.prod(1).cumsum(0)[:-1]])
        return [reference_points(ref_shapes), ss, lsi]

    return pack(grid, pyr), pack(pyr, grid)


def dwconv(sd, prefix, x, H, W):
    B, N, C = x.shape
    n = 18 * 18
    w = sd[prefix + "dwconv.weight"]
    b = sd[prefix + "dwconv.bias"]
    parts = [(x[:, : N - 5 * n], 2 * H + 1, 2 * W + 1), (x[:, N - 5 * n: N - n], H, W), (x[:, N - n:], H // 2, W // 2)]
    outs = []
    for t, hh, ww in parts:
        img = t.transpose(1, 2).reshape(B, C, hh, ww)
        img = F.conv2d(img, w, b, stride=1, padding=1, groups=C)
        outs.append(img.flatten(2).transpose(1, 2))
    return torch.cat(outs, 1)


def conv_ffn(sd, prefix, x, H, W):
    h = F.linear(x, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"])
    h = dwconv(sd, prefix + "dwconv.", h, H, W)
    h = layers.gelu_erf(h)
    return F.linear(h, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])


def _cfg(sd, prefix, n_levels, n_points):
    n_heads = sd[prefix + "attn.attention_weights.weight"].shape[0] // (n_levels * n_points)
    return n_heads


def cavit(sd, prefix, query, ref, feat, spatial_shapes, n_levels, n_points=4, core=msda_core):
    n_heads = _cfg(sd, prefix, n_levels, n_points)
    a = msda_module(sd, prefix + "attn.", layers.layer_norm(sd, prefix + "query_norm.", query), ref,
                    layers.layer_norm(sd, prefix + "feat_norm.", feat), spatial_shapes,
                    n_heads, n_levels, n_points, core=core)
    return query + sd[prefix + "gamma"] * a


def cacnn(sd, prefix, query, ref, feat, spatial_shapes, H, W, n_levels, n_points=4, core=msda_core):
    n_heads = _cfg(sd, prefix, n_levels, n_points)
    a = msda_module(sd, prefix + "attn.", layers.layer_norm(sd, prefix + "query_norm.", query), ref,
                    layers.layer_norm(sd, prefix + "feat_norm.", feat), spatial_shapes,
                    n_heads, n_levels, n_points, core=core)
    query = query + a
    if prefix + "ffn.fc1.weight" in sd:
        query = query + conv_ffn(sd, prefix + "ffn.", layers.layer_norm(sd, prefix + "ffn_norm.", query), H, W)
    return query
