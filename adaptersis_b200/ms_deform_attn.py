"""MSDeformAttn with the reference's constructor, parameters, forward signature, errors and
state_dict keys (backbones/ops/modules/ms_deform_attn.py:63-185); sampling, softmax/location
arithmetic and the four projections run on libasis_b200 kernels."""
import math
import warnings

import torch
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from . import functional as Fn
from .functional import MSDeformAttnFunction, ms_deform_attn_core  # noqa: F401  (re-exported, reference names)


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, ratio=1.0):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        _d_per_head = d_model // n_heads
        if not _is_power_of_2(_d_per_head):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head a "
                          "power of 2 which is more efficient in our CUDA implementation.")
        self.im2col_step = 64
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points
        self.ratio = ratio
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, int(d_model * ratio))
        self.output_proj = nn.Linear(int(d_model * ratio), d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # directional bias grid: one unit-infinity-norm direction per head, scaled by (point index + 1)
        constant_(self.sampling_offsets.weight.data, 0.0)
        theta = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        d = torch.stack([theta.cos(), theta.sin()], -1)
        d = (d / d.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        d = d * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(d.reshape(-1))
        constant_(self.attention_weights.weight.data, 0.0)
        constant_(self.attention_weights.bias.data, 0.0)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.0)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.0)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, gamma=None, residual=None):
        """Reference signature (:120-128).  ``gamma``/``residual`` (ours, optional) fuse the
        caller's ``residual + gamma * output`` into the output projection's epilogue."""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        if not hasattr(self, "_shape_ok") or self._shape_ok[0] is not input_spatial_shapes or self._shape_ok[1] != Len_in:
            # one device->host read per distinct spatial_shapes tensor (they are static per resolution)
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
            self._shape_ok = (input_spatial_shapes, Len_in)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(
                "Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        M, L, P = self.n_heads, self.n_levels, self.n_points
        mode = Fn.get_precision()
        vdt = torch.bfloat16 if mode == "bf16" else torch.float32
        value = Fn.linear(input_flatten, self.value_proj.weight, self.value_proj.bias, out_dtype=vdt)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, M, int(self.ratio * self.d_model) // M)
        offsets = Fn.linear(query, self.sampling_offsets.weight, self.sampling_offsets.bias, out_dtype=torch.float32)
        logits = Fn.linear(query, self.attention_weights.weight, self.attention_weights.bias, out_dtype=torch.float32)
        loc, attn = Fn.MSDAPrepFunction.apply(offsets, logits, reference_points, input_spatial_shapes, M, L, P)
        output = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index, loc, attn,
                                            self.im2col_step)
        if residual is not None:
            return Fn.linear(output, self.output_proj.weight, self.output_proj.bias, gamma=gamma, residual=residual)
        return Fn.linear(output, self.output_proj.weight, self.output_proj.bias, out_dtype=torch.float32)
