"""Autograd nodes of the convolutional stages either side of the hot path -- the spatial prior module
(backbones/encoders.py:4-74) and FeatureDecoder (backbones/decoders.py:92-164) -- on libasis_b200 kernels,
channels-last ([B, H, W, C] = token-major) throughout, so the maps connect to the token tensors of the adapters
without a single transpose:

  Conv2dFunction        nn.Conv2d = asis_im2col + asis_gemm (tcgen05 in bf16 mode; FFMA in fp32 parity mode);
                        backward: transposed GEMM + asis_col2im (gather, no atomics), MN-major GEMM for the weights
  Conv3x3PaddedFunction the 3x3 / stride 1 / pad 1 layers in bf16 mode (two of the stem, four of the decoder): implicit
                        GEMM over zero-padded maps (asis_conv3x3s1_gemm), forward, input and weight gradient; the
                        BatchNorm / resize nodes either side read and write the padded storage directly
  BatchNormFunction     nn.BatchNorm2d / nn.SyncBatchNorm in training mode (+ the ReLU that follows): shifted
                        per-channel sums with fixed-order partials, one fused normalise pass, the same two the other way;
                        SyncBatchNorm exchanges ONE small tensor per layer and direction (statistics all_gather forward,
                        sum all_reduce backward), exactly the reference's semantics (global-batch statistics)
  MaxPoolFunction       nn.MaxPool2d(3, 2, 1) with ATen's tie rule (first maximum in window order)
  SegHeadFunction       the last x2 resize + the 64 -> n_classes 3x3 head, fused: contraction at low resolution
"""
import weakref

import torch
import torch.distributed as dist
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import functional as Fn
from . import kernels as K
from ._lib import MAJOR_K, MAJOR_MN


def _conv_weight(weight, cdt, ldk):
    """[Cout, Cin, k, k] parameter -> [Cout, ldk] in the compute dtype, K = (ky, kx, cin); cached per version."""
    key = ("cw", id(weight), cdt)
    hit = Fn._wcache.get(key)
    ver = (weight._version, weight.data_ptr())
    if hit is not None and hit[0]() is weight and hit[1] == ver:
        return hit[2]
    Cout = weight.shape[0]
    w2 = weight.detach().permute(0, 2, 3, 1).reshape(Cout, -1)
    if w2.shape[1] != ldk:
        w2 = torch.nn.functional.pad(w2, (0, ldk - w2.shape[1]))
    w2 = K.cast(w2.contiguous().float(), cdt)
    if isinstance(weight, torch.nn.Parameter):
        if hit is None:
            weakref.finalize(weight, Fn._wcache.pop, key, None)
        Fn._wcache[key] = (weakref.ref(weight), ver, w2)
    return w2


class Conv2dFunction(Function):
    """x [B, H, W, Cin] channels-last -> [B, Ho, Wo, Cout]; weight in nn.Conv2d's own layout [Cout, Cin, k, k]."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, mode, out_dtype):
        comp, cdt = Fn._cfg(mode)
        B, H, W, Cin = x.shape
        Cout, _, k, _ = weight.shape
        cols, (Ho, Wo) = K.im2col(x, k, stride, pad, cdt)
        R, ldk = cols.shape
        w2 = _conv_weight(weight, cdt, ldk)
        y, _ = K.gemm(comp, cols, MAJOR_K, w2, MAJOR_K, R, Cout, ldk, out_dtype or cdt, bias=Fn._f32(bias))
        ctx.save_for_backward(cols, weight)
        ctx.meta = (B, H, W, Cin, Cout, k, stride, pad, Ho, Wo, mode, bias is not None, x.dtype)
        return y.view(B, Ho, Wo, Cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        cols, weight = ctx.saved_tensors
        B, H, W, Cin, Cout, k, stride, pad, Ho, Wo, mode, has_bias, xdt = ctx.meta
        comp, cdt = Fn._cfg(mode)
        R, ldk = cols.shape
        Kd = k * k * Cin
        dy2 = K.cast(dy.reshape(R, Cout), cdt)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            w2 = _conv_weight(weight, cdt, ldk)
            dcols, _ = K.gemm(comp, dy2, MAJOR_K, w2, MAJOR_MN, R, ldk, Cout, cdt)
            dx = K.col2im(dcols, B, H, W, Cin, k, stride, pad)
            if dx.dtype != xdt:
                dx = dx.to(xdt)
        if ctx.needs_input_grad[1]:
            dw2, _ = K.gemm(comp, dy2, MAJOR_MN, cols, MAJOR_MN, Cout, ldk, R, torch.float32)
            dw = dw2[:, :Kd].reshape(Cout, k, k, Cin).permute(0, 3, 1, 2).to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(dy2)
        return dx, dw, db, None, None, None, None


def conv2d(x, weight, bias, stride, pad, out_dtype=None, mode=None):
    mode = mode or Fn.get_precision()
    # the bf16 kernels move 8-channel (16-byte) vectors and TMA needs 16-byte row pitches: layers whose channel
    # counts are not multiples of 8 (toy configurations only; the real ones are multiples of 64) run in fp32
    Cout, Cin = weight.shape[0], weight.shape[1]
    if mode == "bf16" and (Cout % 8 or (Cin % 8 and x.requires_grad)):
        mode = "fp32"
    return Conv2dFunction.apply(x, weight, bias, stride, pad, mode, out_dtype)


def implicit_ok(conv, mode=None):
    """can this nn.Conv2d run as the implicit GEMM (asis_conv3x3s1_gemm: 3x3 / stride 1 / pad 1, channels in 64s, bf16)?"""
    mode = mode or Fn.get_precision()
    return (mode == "bf16" and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (1, 1)
            and conv.groups == 1 and conv.in_channels % 64 == 0 and conv.out_channels % 64 == 0)


class Conv3x3PaddedFunction(Function):
    """3x3 / stride 1 / pad 1 convolution over a ZERO-PADDED channels-last bf16 map, as an implicit GEMM on the tcgen05
    kernel (tap (ky, kx) of the K loop shifts the TMA row coordinate: no column matrix in memory).
    xp [B, H+2, W+2, Cin] (zero border) -> yp [B, H+2, W+2, Cout] whose border rows are NOT meaningful: the consumer
    (BatchNormFunction with pad_in=1) reads the logical pixels only.  Backward contract: the incoming gradient is in the
    same padded storage with a ZERO border (BatchNormFunction writes it so); the returned dxp has a meaningless border."""

    @staticmethod
    def forward(ctx, xp, weight, bias, out_dtype):
        B, Hp, Wp, Cin = xp.shape
        Cout = weight.shape[0]
        assert xp.dtype == torch.bfloat16 and xp.is_contiguous()
        w2 = _conv_weight(weight, torch.bfloat16, 9 * Cin)
        yp = torch.empty(B, Hp, Wp, Cout, dtype=out_dtype or torch.bfloat16, device=xp.device)
        K.conv3x3s1_gemm(0, xp, w2, yp, Fn._f32(bias), B, Hp - 2, Wp - 2, Cin, Cout)
        ctx.save_for_backward(xp, weight)
        ctx.has_bias = bias is not None
        return yp

    @staticmethod
    @once_differentiable
    def backward(ctx, dyp):
        xp, weight = ctx.saved_tensors
        B, Hp, Wp, Cin = xp.shape
        Cout = weight.shape[0]
        dyp = K.cast(dyp.contiguous(), torch.bfloat16)
        dxp = dw = db = None
        if ctx.needs_input_grad[0]:
            w2 = _conv_weight(weight, torch.bfloat16, 9 * Cin)
            dxp = torch.empty_like(xp)
            K.conv3x3s1_gemm(1, dyp, w2, dxp, None, B, Hp - 2, Wp - 2, Cin, Cout)
        if ctx.needs_input_grad[1]:
            dw2 = torch.empty(Cout, 9 * Cin, dtype=torch.float32, device=xp.device)
            K.conv3x3s1_gemm(2, dyp, xp, dw2, None, B, Hp - 2, Wp - 2, Cin, Cout)
            dw = dw2.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(dyp.view(-1, Cout))              # the border is zero: all rows may be summed
        return dxp, dw, db, None


def conv3x3_padded(xp, weight, bias, out_dtype=None):
    return Conv3x3PaddedFunction.apply(xp, weight, bias, out_dtype)


class RepadFunction(Function):
    """change the storage padding (and dtype) of a channels-last map: [B, H+2pi, W+2pi, C] -> [B, H+2po, W+2po, C] with a
    zero border (asis_bn_apply with a = 1, b = 0)."""

    @staticmethod
    def forward(ctx, x, pad_in, pad_out, dtype):
        C = x.shape[-1]
        one, zero = torch.ones(C, device=x.device), torch.zeros(C, device=x.device)
        ctx.meta = (pad_in, pad_out, x.dtype)
        ctx.save_for_backward(one, zero)
        return K.bn_apply(x, one, zero, zero, False, dtype or x.dtype, pad_in, pad_out)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        one, zero = ctx.saved_tensors
        pad_in, pad_out, xdt = ctx.meta
        return K.cast(K.bn_apply(dy.contiguous(), one, zero, zero, False, dy.dtype, pad_out, pad_in), xdt), None, None, None


def repad(x, pad_in, pad_out, dtype=None):
    return RepadFunction.apply(x, pad_in, pad_out, dtype)


def _sync_world(bn):
    """process group size a SyncBatchNorm layer synchronises over (1: plain batch statistics)."""
    if not isinstance(bn, torch.nn.SyncBatchNorm) or not dist.is_available() or not dist.is_initialized():
        return 1, None
    group = bn.process_group
    return dist.get_world_size(group), group


class BatchNormFunction(Function):
    """BatchNorm over a channels-last map, optionally followed by ReLU, as ONE node.
    Training mode: statistics of the (global, for SyncBatchNorm) batch; eval mode: the running statistics."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn, relu, out_dtype, pad_in=0, pad_out=0):
        B, Hs, Ws, C = x.shape
        H, W = Hs - 2 * pad_in, Ws - 2 * pad_in
        n_local = B * H * W
        world, group = _sync_world(bn)
        training = bn.training or bn.running_mean is None
        w32, b32 = Fn._f32(weight), Fn._f32(bias)
        if training:
            shift = x[0, pad_in, pad_in].float().contiguous()       # any value near the data: kills the cancellation
            s = K.chan_stats(x, shift, pad_in)
            mean = shift + s[0] / n_local
            m2 = s[1] - s[0] * s[0] / n_local                       # sum of squared deviations from the local mean
            n_tot = n_local
            if world > 1:
                # ONE all_gather per layer: [mean, m2, count] of every rank, merged with Chan's formula
                # (torch.full: a fill kernel -- a host tensor copied in would break CUDA-graph capture of the step)
                mine = torch.cat([mean, m2, torch.full((1,), float(n_local), dtype=mean.dtype, device=mean.device)])
                every = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(every, mine, group=group)
                st = torch.stack(every)
                cnt = st[:, 2 * C]
                n_tot = cnt.sum()                                   # stays on the device: no host read in the step
                mean_r, m2_r = st[:, :C], st[:, C:2 * C]
                mean = (mean_r * cnt[:, None]).sum(0) / n_tot
                m2 = (m2_r + cnt[:, None] * (mean_r - mean) ** 2).sum(0)
            var = m2 / n_tot
            if bn.running_mean is not None and bn.track_running_stats:
                with torch.no_grad():
                    mom = bn.momentum if bn.momentum is not None else 0.1
                    unbiased = m2 / (n_tot - 1).clamp_min(1.0) if torch.is_tensor(n_tot) else m2 / max(n_tot - 1, 1)
                    bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
                    bn.running_var.mul_(1 - mom).add_(unbiased, alpha=mom)
                    bn.num_batches_tracked += 1
        else:
            mean, var, n_tot = bn.running_mean.float(), bn.running_var.float(), n_local
        rstd = torch.rsqrt(var + bn.eps)
        a = (w32 * rstd).contiguous()
        b, mean = b32.contiguous(), mean.contiguous()           # y = a (x - mean) + b: the centred form (csrc/conv.cu)
        y = K.bn_apply(x, a, b, mean, relu, out_dtype or x.dtype, pad_in, pad_out)
        ctx.save_for_backward(x, a, b, mean, rstd.contiguous())
        ctx.meta = (relu, training, world, group, n_tot, pad_in, pad_out)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, a, b, mean, rstd = ctx.saved_tensors
        relu, training, world, group, n_tot, pad_in, pad_out = ctx.meta
        dy = dy if dy.dtype == x.dtype else dy.to(x.dtype)
        s = K.chan_stats_backward(x, dy, a, b, mean, rstd, relu, pad_in, pad_out)
        dweight, dbias = s[1].clone(), s[0].clone()       # local sums: the gradient all-reduce averages them (as DDP does)
        dx = None
        if ctx.needs_input_grad[0]:
            if training:
                if world > 1:
                    dist.all_reduce(s, group=group)         # ONE all_reduce per layer: [sum dz, sum dz * xhat]
                c1, c2 = (s[0] / n_tot).contiguous(), (s[1] / n_tot).contiguous()
            else:
                c1 = c2 = torch.zeros_like(mean)
            dx = K.bn_apply_backward(x, dy, a, b, mean, rstd, c1, c2, relu, pad_in, pad_out)
        return dx, dweight, dbias, None, None, None, None, None


def batch_norm(x, bn, relu=True, out_dtype=None, pad_in=0, pad_out=0):
    """`bn`: the nn.BatchNorm2d / nn.SyncBatchNorm module that holds parameters and running statistics.
    x is stored with `pad_in` border pixels (not part of the statistics), the result with `pad_out` (zero) border pixels;
    the gradient comes back in the latter storage and leaves in the former (zero border)."""
    return BatchNormFunction.apply(x, bn.weight, bn.bias, bn, relu, out_dtype, pad_in, pad_out)


def conv_bn_relu(x, conv, bn, x_pad=0, out_pad=0, record=None, name=None):
    """relu(bn(conv(x))) on channels-last maps; x_pad = 1: x is zero-padded storage and `conv` runs as the implicit GEMM.
    ``record`` (parity tests): dict that receives the ReLU mask of the layer as ``record[name]`` [B, C, H, W] bool."""
    if x_pad:
        y = conv3x3_padded(x, conv.weight, conv.bias)
    else:
        y = conv2d(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0])
    y = batch_norm(y, bn, relu=True, pad_in=x_pad, pad_out=out_pad)
    if record is not None:
        z = y.detach()
        z = z[:, out_pad:z.shape[1] - out_pad, out_pad:z.shape[2] - out_pad]
        record[name] = (z > 0).permute(0, 3, 1, 2)
    return y


class MaxPoolFunction(Function):
    @staticmethod
    def forward(ctx, x):
        y, idx = K.maxpool3x3s2_forward(x)
        ctx.save_for_backward(idx)
        ctx.hw = (x.shape[1], x.shape[2])
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (idx,) = ctx.saved_tensors
        return K.maxpool3x3s2_backward(gy, idx, *ctx.hw)


def maxpool3x3s2(x):
    return MaxPoolFunction.apply(x)


class SegHeadFunction(Function):
    """nn.Upsample(x2, bilinear, align_corners=True) -> 3x3 / pad 1 convolution to <= 4 classes, fused
    (decoders.py:125-129): z [B,H,W,C] channels-last -> logits [B,2H,2W,CO] f32.  The channel contraction runs at the
    low resolution (csrc/conv.cu: the upsampled C-channel map is never materialised)."""

    @staticmethod
    def forward(ctx, z, weight, bias):
        CO, C = weight.shape[0], weight.shape[1]
        w2 = weight.detach().float().permute(2, 3, 0, 1).reshape(9 * CO, C).contiguous()     # row (ky*3 + kx)*CO + co
        y = K.seg_head_forward(z, w2, Fn._f32(bias), CO)
        ctx.save_for_backward(z, w2)
        ctx.meta = (CO, C, weight.dtype, bias is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        z, w2 = ctx.saved_tensors
        CO, C, wdt, has_bias = ctx.meta
        gz, gw2 = K.seg_head_backward(z, w2, gy, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        gw = gw2.view(3, 3, CO, C).permute(2, 3, 0, 1).to(wdt) if gw2 is not None else None
        gb = gy.reshape(-1, CO).float().sum(0) if (has_bias and ctx.needs_input_grad[2]) else None        # [CO] <= 4
        return gz, gw, gb


def seg_head(z, weight, bias):
    return SegHeadFunction.apply(z, weight, bias)
