"""Autograd nodes of the convolutional stages either side of the hot path -- the spatial prior module
(backbones/encoders.py:4-74) and FeatureDecoder (backbones/decoders.py:92-164) -- on libasis_b200 kernels,
channels-last ([B, H, W, C] = token-major) throughout, so the maps connect to the token tensors of the adapters
without a single transpose:

  Conv2dFunction        nn.Conv2d = asis_im2col + asis_gemm (tcgen05 in bf16 mode; FFMA in fp32 parity mode);
                        backward: transposed GEMM + asis_col2im (gather, no atomics), MN-major GEMM for the weights
  BatchNormFunction     nn.BatchNorm2d / nn.SyncBatchNorm in training mode (+ the ReLU that follows): shifted
                        per-channel sums with fixed-order partials, one fused normalise pass, the same two the other way;
                        SyncBatchNorm exchanges ONE small tensor per layer and direction (statistics all_gather forward,
                        sum all_reduce backward), exactly the reference's semantics (global-batch statistics)
  MaxPoolFunction       nn.MaxPool2d(3, 2, 1) with ATen's tie rule (first maximum in window order)
  SmallConvFunction     the 64 -> n_classes 3x3 head: direct kernels (its column matrix would be 3.7 G elements)
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
    return Conv2dFunction.apply(x, weight, bias, stride, pad, mode or Fn.get_precision(), out_dtype)


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
    def forward(ctx, x, weight, bias, bn, relu, out_dtype):
        B, H, W, C = x.shape
        n_local = B * H * W
        world, group = _sync_world(bn)
        training = bn.training or bn.running_mean is None
        w32, b32 = Fn._f32(weight), Fn._f32(bias)
        if training:
            shift = x[0, 0, 0].float().contiguous()                 # any value near the data: kills the cancellation
            s = K.chan_stats(x, shift)
            mean = shift + s[0] / n_local
            m2 = s[1] - s[0] * s[0] / n_local                       # sum of squared deviations from the local mean
            n_tot = n_local
            if world > 1:
                # ONE all_gather per layer: [mean, m2, count] of every rank, merged with Chan's formula
                mine = torch.cat([mean, m2, mean.new_tensor([float(n_local)])])
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
        b = (b32 - mean * a).contiguous()
        y = K.bn_apply(x, a, b, relu, out_dtype or x.dtype)
        ctx.save_for_backward(x, a, b, mean.contiguous(), rstd.contiguous())
        ctx.meta = (relu, training, world, group, n_tot)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, a, b, mean, rstd = ctx.saved_tensors
        relu, training, world, group, n_tot = ctx.meta
        dy = dy if dy.dtype == x.dtype else dy.to(x.dtype)
        s = K.chan_stats_backward(x, dy, a, b, mean, rstd, relu)
        dweight, dbias = s[1].clone(), s[0].clone()       # local sums: the gradient all-reduce averages them (as DDP does)
        dx = None
        if ctx.needs_input_grad[0]:
            if training:
                if world > 1:
                    dist.all_reduce(s, group=group)         # ONE all_reduce per layer: [sum dz, sum dz * xhat]
                c1, c2 = (s[0] / n_tot).contiguous(), (s[1] / n_tot).contiguous()
            else:
                c1 = c2 = torch.zeros_like(mean)
            dx = K.bn_apply_backward(x, dy, a, b, mean, rstd, c1, c2, relu)
        return dx, dweight, dbias, None, None, None


def batch_norm(x, bn, relu=True, out_dtype=None):
    """`bn`: the nn.BatchNorm2d / nn.SyncBatchNorm module that holds parameters and running statistics."""
    return BatchNormFunction.apply(x, bn.weight, bn.bias, bn, relu, out_dtype)


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


class SmallConvFunction(Function):
    """3x3 / stride 1 / pad 1 convolution to a handful of channels: x [B,H,W,C] -> [B,H,W,CO] f32."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w = weight.detach().permute(0, 2, 3, 1).contiguous().float()          # [CO, 3, 3, C]
        y = K.smallconv3x3_forward(x, w, Fn._f32(bias))
        ctx.save_for_backward(x, w)
        ctx.wshape = (weight.shape, weight.dtype, bias is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        wshape, wdt, has_bias = ctx.wshape
        gx, gw = K.smallconv3x3_backward(x, w, gy, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        if gw is not None:
            gw = gw.permute(0, 3, 1, 2).to(wdt)
        gb = gy.reshape(-1, gy.shape[-1]).float().sum(0) if (has_bias and ctx.needs_input_grad[2]) else None     # [CO] <= 4
        return gx, gw, gb


def smallconv3x3(x, weight, bias):
    return SmallConvFunction.apply(x, weight, bias)
