"""Oracle: multi-scale deformable attention (test infrastructure only).

Restates the mathematics of the reference's
``backbones/ops/modules/ms_deform_attn.py``:

  * ``msda_core``      <- ``ms_deform_attn_core_pytorch``  (:33-54)
  * ``msda_locations`` <- location arithmetic in ``MSDeformAttn.forward`` (:161-175)
  * ``msda_module``    <- ``MSDeformAttn.forward`` (:120-185)
  * ``msda_reset_bias``<- directional bias grid of ``_reset_parameters`` (:101-112)

The reference goes through ``F.grid_sample``; here the bilinear gather is
written out explicitly (floor / 4 corners / zero padding) so that the oracle
is an independent formulation, differentiable by plain autograd in fp32 or
fp64.  Pixel coordinate convention (reference :39,47-49 with
``align_corners=False``): ``x_pix = loc_x * W - 0.5`` evaluated the way ATen
evaluates it, ``((2*loc-1) + 1) * W - 1) / 2``.
"""
import math

import torch
import torch.nn.functional as F


def _shapes_list(spatial_shapes):
    if torch.is_tensor(spatial_shapes):
        return [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    return [(int(h), int(w)) for h, w in spatial_shapes]


def msda_core(value, spatial_shapes, sampling_locations, attention_weights):
    """value [N,S,M,D]; sampling_locations [N,Lq,M,L,P,2] (x,y in [0,1]);
    attention_weights [N,Lq,M,L,P]  ->  [N,Lq,M*D]."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    shapes = _shapes_list(spatial_shapes)
    assert sum(h * w for h, w in shapes) == S
    out = value.new_zeros(N, Lq, M, D)
    start = 0
    n_idx = torch.arange(N, device=value.device).view(N, 1, 1, 1)
    m_idx = torch.arange(M, device=value.device).view(1, 1, M, 1)
    for lvl, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W]                     # [N,HW,M,D]
        loc = sampling_locations[:, :, :, lvl]                # [N,Lq,M,P,2]
        gx = 2 * loc[..., 0] - 1
        gy = 2 * loc[..., 1] - 1
        x = ((gx + 1) * W - 1) / 2
        y = ((gy + 1) * H - 1) / 2
        x0 = torch.floor(x)
        y0 = torch.floor(y)
        lx = x - x0
        ly = y - y0
        x0 = x0.long()
        y0 = y0.long()
        aw = attention_weights[:, :, :, lvl]                  # [N,Lq,M,P]
        for dy in (0, 1):
            for dx in (0, 1):
                xi = x0 + dx
                yi = y0 + dy
                wgt = (lx if dx else 1 - lx) * (ly if dy else 1 - ly)
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                pix = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1))   # [N,Lq,M,P]
                g = v[n_idx, pix, m_idx]                               # [N,Lq,M,P,D]
                coef = (wgt * aw * ok.to(value.dtype)).unsqueeze(-1)
                out = out + (g * coef).sum(3)
        start += H * W
    return out.reshape(N, Lq, M * D)


def msda_locations(reference_points, sampling_offsets, spatial_shapes, n_points):
    """reference_points [N|1,Lq,L|1,2or4], sampling_offsets [N,Lq,M,L,P,2]."""
    shapes = spatial_shapes if torch.is_tensor(spatial_shapes) else torch.as_tensor(spatial_shapes)
    shapes = shapes.to(sampling_offsets.device)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1).to(sampling_offsets.dtype)  # (W,H)
        return reference_points[:, :, None, :, None, :] + sampling_offsets / norm[None, None, None, :, None, :]
    if reference_points.shape[-1] == 4:
        return (reference_points[:, :, None, :, None, :2]
                + sampling_offsets / n_points * reference_points[:, :, None, :, None, 2:] * 0.5)
    raise ValueError(
        "Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))


def msda_module(sd, prefix, query, reference_points, input_flatten, spatial_shapes,
                n_heads, n_levels, n_points, padding_mask=None, core=msda_core):
    """Functional MSDeformAttn.forward over a reference state_dict."""
    N, Lq, C = query.shape
    _, S, _ = input_flatten.shape
    shapes = _shapes_list(spatial_shapes)
    assert sum(h * w for h, w in shapes) == S
    value = F.linear(input_flatten, sd[prefix + "value_proj.weight"], sd[prefix + "value_proj.bias"])
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)
    Cv = value.shape[-1]
    value = value.view(N, S, n_heads, Cv // n_heads)
    off = F.linear(query, sd[prefix + "sampling_offsets.weight"], sd[prefix + "sampling_offsets.bias"])
    off = off.view(N, Lq, n_heads, n_levels, n_points, 2)
    aw = F.linear(query, sd[prefix + "attention_weights.weight"], sd[prefix + "attention_weights.bias"])
    aw = torch.softmax(aw.view(N, Lq, n_heads, n_levels * n_points), -1)
    aw = aw.view(N, Lq, n_heads, n_levels, n_points)
    loc = msda_locations(reference_points, off, torch.as_tensor(shapes), n_points)
    out = core(value, shapes, loc, aw)
    return F.linear(out, sd[prefix + "output_proj.weight"], sd[prefix + "output_proj.bias"])


def msda_reset_bias(n_heads, n_levels, n_points):
    """The sampling_offsets.bias the reference constructor produces (:101-112):
    per head a unit-infinity-norm direction, scaled by (point index + 1)."""
    theta = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
    d = torch.stack([theta.cos(), theta.sin()], -1)
    d = d / d.abs().max(-1, keepdim=True)[0]
    g = d.view(n_heads, 1, 1, 2).repeat(1, n_levels, n_points, 1)
    scale = torch.arange(1, n_points + 1, dtype=torch.float32).view(1, 1, n_points, 1)
    return (g * scale).reshape(-1)
