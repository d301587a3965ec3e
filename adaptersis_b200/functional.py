"""Autograd nodes of the hot path.  Every node's forward and backward is a fixed sequence of
libasis_b200 kernels (see kernels.py); torch only allocates tensors and records the graph.

Precision modes (BASELINE.json north_star):
  "fp32": FFMA GEMM / attention kernels, fp32 activations            -- parity mode, 1e-4
  "bf16": tcgen05 GEMM / flash attention, bf16 operands + activations,
          fp32 accumulation, fp32 residual stream and LayerNorm stats -- performance mode, 2e-2
"""
import contextlib
import os
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels as K
from ._lib import (BF16, EPI_ACCUMULATE, EPI_DGELU, EPI_GELU, EPI_GELU_GRAD, EPI_MUL_AUX, EPI_NONE, EPI_SCALE_RESIDUAL, F32,
                   MAJOR_K, MAJOR_MN)

_MODE = [os.environ.get("ASIS_PRECISION", "fp32")]


def get_precision():
    return _MODE[0]


def set_precision(mode):
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _MODE[0] = mode
    # (no library convolution is left on the path; kept so that user code mixing in cuDNN layers stays out of TF32
    # in parity mode)
    torch.backends.cudnn.allow_tf32 = mode != "fp32"


@contextlib.contextmanager
def precision(mode):
    old = _MODE[0]
    set_precision(mode)
    try:
        yield
    finally:
        set_precision(old)


set_precision(_MODE[0])


def _cfg(mode=None):
    mode = mode or _MODE[0]
    return (BF16, torch.bfloat16) if mode == "bf16" else (F32, torch.float32)


# bf16 copies of fp32 master weights.  One entry per live Parameter, keyed by id() and evicted by a
# weakref finalizer when the Parameter dies (no strong reference is held, rebuilt models do not leak).
# An entry is valid while (version counter, storage pointer) are unchanged: every in-place update that
# autograd can see -- optimizer.step(), load_state_dict(), p.copy_()/p.mul_() under no_grad -- bumps
# the version.  Writes through ``p.data`` do NOT (``.data`` is a detached alias with its own counter):
# call ``invalidate_weight_cache()`` after such writes, or run with ASIS_CHECK_WEIGHT_CACHE=1, which
# re-casts on every use and raises if a cached copy went stale.
_wcache = {}
_CHECK_WCACHE = bool(int(os.environ.get("ASIS_CHECK_WEIGHT_CACHE", "0")))


def invalidate_weight_cache():
    """Drop every cached low-precision weight copy (needed only after writes through ``param.data``)."""
    for key in [k for k in _wcache if not (isinstance(k, tuple) and k[0] == "ones")]:
        _wcache.pop(key, None)


def _operand(t, tdtype):
    """Weight (or any tensor) in the compute dtype; parameters are cached per version."""
    if t.dtype == tdtype:
        return t.detach()
    if not isinstance(t, torch.nn.Parameter):
        return K.cast(t.detach(), tdtype)
    key = id(t)
    hit = _wcache.get(key)
    if hit is not None and hit[0]() is t and hit[1] == (t._version, t.data_ptr()) and hit[2].dtype == tdtype:
        if _CHECK_WCACHE and not torch.equal(hit[2], K.cast(t.detach(), tdtype)):
            raise RuntimeError("stale low-precision weight copy: the parameter was modified through .data "
                               "(no version bump); call adaptersis_b200.invalidate_weight_cache() after such writes")
        return hit[2]
    c = K.cast(t.detach(), tdtype)
    if hit is None:
        weakref.finalize(t, _wcache.pop, key, None)
    _wcache[key] = (weakref.ref(t), (t._version, t.data_ptr()), c)
    return c


def _f32(t):
    return None if t is None else (t.detach() if t.dtype == torch.float32 else t.detach().float())


def _ones(n, device):
    key = ("ones", n, device)
    hit = _wcache.get(key)
    if hit is None:
        hit = torch.ones(n, dtype=torch.float32, device=device)
        _wcache[key] = hit
    return hit


# ------------------------------------------------------------------------------------------------
class MSDeformAttnFunction(Function):
    """Drop-in for the reference's MSDeformAttnFunction (ms_deform_attn.py:17-30) -- same apply()
    signature -- with the backward the reference lacks."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step=64):
        out = K.msda_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                             attention_weights)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, ss, lsi, loc, aw = ctx.saved_tensors
        gv, gl, ga = K.msda_backward(value, ss, lsi, loc, aw, grad_output)
        return gv, None, None, gl.to(loc.dtype), ga.to(aw.dtype), None


def ms_deform_attn_core(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Same signature as the reference's ms_deform_attn_core_pytorch (:33-54), CUDA kernels behind."""
    ss = torch.as_tensor(value_spatial_shapes, dtype=torch.long, device=value.device)
    lsi = torch.cat([ss.new_zeros(1), ss.prod(1).cumsum(0)[:-1]])
    return MSDeformAttnFunction.apply(value, ss, lsi, sampling_locations, attention_weights, 64)


class MSDAPrepFunction(Function):
    """softmax over L*P + sampling-location arithmetic (ms_deform_attn.py:156-171), fused."""

    @staticmethod
    def forward(ctx, offsets, logits, reference_points, spatial_shapes, M, L, P):
        N, Lq = offsets.shape[0], offsets.shape[1]
        loc, attn = K.msda_prep_forward(offsets, logits, reference_points, spatial_shapes, N, Lq, M, L, P)
        ctx.save_for_backward(attn, reference_points, spatial_shapes)
        ctx.dims = (N, Lq, M, L, P, offsets.dtype)
        return loc, attn

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loc, grad_attn):
        attn, ref, ss = ctx.saved_tensors
        N, Lq, M, L, P, odt = ctx.dims
        goff, glog = K.msda_prep_backward(grad_loc.float(), grad_attn.float(), attn, ref, ss, odt, N, Lq, M, L, P)
        return goff, glog, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
class LayerNormFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if x2.dtype not in (torch.float32, torch.bfloat16):
            x2 = x2.float()
        y, mean, rstd = K.layernorm_forward(x2, _f32(weight), _f32(bias), eps, out_dtype)
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.shp = shp
        return y.view(shp)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        dy2 = dy.reshape(x2.shape)
        if dy2.dtype not in (torch.float32, torch.bfloat16):
            dy2 = dy2.float()
        want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dw, db = K.layernorm_backward(dy2, x2, _f32(weight), mean, rstd, None, want)
        return dx.view(ctx.shp), dw, db, None, None


def layer_norm(x, weight, bias, eps=1e-6, out_dtype=None):
    return LayerNormFunction.apply(x, weight, bias, eps, out_dtype or _cfg()[1])


class LayerNormForkFunction(Function):
    """(LN(x), x): the normalised tensor plus the input handed on as a second output, for the places where a tensor
    feeds a LayerNorm AND continues as a residual / a later operand (adapter_blocks.py:127-143, :171-181).  As two
    separate uses autograd sums the two gradients with an elementwise pass over [B, 6949, C] f32 (12 ATen adds,
    1.5 ms per step); as one node the LayerNorm backward kernel adds the pass-through gradient while it writes dx
    (its ``dres`` operand, the same fusion BlockFunction uses)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if x2.dtype not in (torch.float32, torch.bfloat16):
            x2 = x2.float()
        y, mean, rstd = K.layernorm_forward(x2, _f32(weight), _f32(bias), eps, out_dtype)
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.shp = shp
        ctx.set_materialize_grads(False)        # an unused output arrives as None, not as a tensor of zeros to add
        return y.view(shp), x           # (x returned as is: autograd hands out an alias attached to this node)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, dpass):
        x2, weight, mean, rstd = ctx.saved_tensors
        if dy is None:
            return dpass, None, None, None, None
        dy2 = dy.reshape(x2.shape)
        if dy2.dtype not in (torch.float32, torch.bfloat16):
            dy2 = dy2.float()
        want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dres = dpass.reshape(x2.shape) if dpass is not None and ctx.needs_input_grad[0] else None
        dx, dw, db = K.layernorm_backward(dy2, x2, _f32(weight), mean, rstd, dres, want)
        return dx.view(ctx.shp), dw, db, None, None


def layer_norm_fork(x, weight, bias, eps=1e-6, out_dtype=None):
    """-> (LayerNorm(x), x): use the second result wherever ``x`` itself is needed afterwards."""
    if not (torch.is_grad_enabled() and x.requires_grad):
        return layer_norm(x, weight, bias, eps, out_dtype), x
    return LayerNormForkFunction.apply(x, weight, bias, eps, out_dtype or _cfg()[1])


# ------------------------------------------------------------------------------------------------
def _linear_backward(comp, cdt, dy2, x2, weight, need_dx, need_dw, need_db, dx_dtype=None, dgelu_aux=None):
    """dy2 [R, N] compute dtype; x2 [R, K] compute dtype; weight [N, K].
    returns dx [R, K] (dx_dtype), dW [N, K] f32, db [N] f32.
    ``dgelu_aux``: the GELU derivative saved by the forward's GELU_GRAD epilogue -- the input gradient is multiplied
    by it in the epilogue of the dgrad GEMM."""
    R, N = dy2.shape
    Kd = x2.shape[1] if x2 is not None else weight.shape[1]
    dx = dw = db = None
    if need_dx:
        w = _operand(weight, cdt)
        epi = EPI_MUL_AUX if dgelu_aux is not None else EPI_NONE
        dx, _ = K.gemm(comp, dy2, MAJOR_K, w, MAJOR_MN, R, Kd, N, dx_dtype or cdt, epilogue=epi, aux=dgelu_aux)
    if need_dw:
        dw, _ = K.gemm(comp, dy2, MAJOR_MN, x2, MAJOR_MN, N, Kd, R, torch.float32)
    if need_db:
        db = K.colsum(dy2)
    return dx, dw, db


class LinearFunction(Function):
    """nn.Linear (+ optional fused  residual + gamma * (.)  epilogue: LayerScale / injector gamma,
    block.py:91-94, adapter_blocks.py:136,176)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, residual, out_dtype, mode):
        comp, cdt = _cfg(mode)
        shp = x.shape
        Kd = shp[-1]
        N = weight.shape[0]
        x2 = K.cast(x.reshape(-1, Kd), cdt) if x.dtype != cdt else x.reshape(-1, Kd)
        R = x2.shape[0]
        w = _operand(weight, cdt)
        fused = residual is not None
        if fused:
            g = _f32(gamma) if gamma is not None else _ones(N, x.device)
            res2 = residual.reshape(-1, N)
            res2 = res2 if res2.dtype == torch.float32 else res2.float()
            need_u = gamma is not None and gamma.requires_grad
            y, u = K.gemm(comp, x2, MAJOR_K, w, MAJOR_K, R, N, Kd, torch.float32, epilogue=EPI_SCALE_RESIDUAL,
                          bias=_f32(bias), gamma=g, residual=res2, want_aux_dtype=cdt if need_u else None)
        else:
            y, u = K.gemm(comp, x2, MAJOR_K, w, MAJOR_K, R, N, Kd, out_dtype or cdt, bias=_f32(bias))
        ctx.save_for_backward(x2, weight, gamma, u)
        ctx.meta = (shp, fused, mode, bias is not None, x.dtype,
                    residual.shape if fused else None, residual.dtype if fused else None)
        return y.view(*shp[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, weight, gamma, u = ctx.saved_tensors
        shp, fused, mode, has_bias, xdt, res_shape, res_dt = ctx.meta
        comp, cdt = _cfg(mode)
        N = weight.shape[0]
        dy2 = dy.reshape(-1, N)
        dgamma = dres = None
        if fused:
            dy2 = dy2 if dy2.dtype == torch.float32 else dy2.float()
            if ctx.needs_input_grad[4]:
                dres = dy2.view(res_shape).to(res_dt)
            if gamma is not None:
                if ctx.needs_input_grad[3]:
                    dgamma = K.colsum(dy2, u)
                dy2 = K.scale_cols(dy2, _f32(gamma), cdt)
            else:
                dy2 = K.cast(dy2, cdt)
        else:
            dy2 = K.cast(dy2, cdt) if dy2.dtype != cdt else dy2
        dx, dw, db = _linear_backward(comp, cdt, dy2, x2, weight, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                      has_bias and ctx.needs_input_grad[2])
        if dx is not None:
            dx = dx.view(shp)
        return dx, dw, db, dgamma, dres, None, None


def linear(x, weight, bias=None, gamma=None, residual=None, out_dtype=None, mode=None):
    return LinearFunction.apply(x, weight, bias, gamma, residual, out_dtype, mode or get_precision())


# ------------------------------------------------------------------------------------------------
class AttentionFunction(Function):
    """softmax((q * hd^-0.5) k^T) v over packed qkv [B, T, 3*H*hd] (attention.py:56-69)."""

    @staticmethod
    def forward(ctx, qkv, num_heads, mode):
        comp, cdt = _cfg(mode)
        B, T, C3 = qkv.shape
        hd = C3 // 3 // num_heads
        q = K.cast(qkv, cdt) if qkv.dtype != cdt else qkv
        out, lse = K.attention_forward(comp, q, B, T, num_heads, hd)
        ctx.save_for_backward(q, out, lse)
        ctx.meta = (B, T, num_heads, hd, mode, qkv.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        q, out, lse = ctx.saved_tensors
        B, T, H, hd, mode, qdt = ctx.meta
        comp, cdt = _cfg(mode)
        dqkv = K.attention_backward(comp, q, out, lse, dout, B, T, H, hd)
        return dqkv.to(qdt), None, None


def attention(qkv, num_heads, mode=None):
    return AttentionFunction.apply(qkv, num_heads, mode or get_precision())


# ------------------------------------------------------------------------------------------------
class MlpFunction(Function):
    """fc2(gelu_erf(fc1 x)) (mlp.py:34-40); GELU fused into the fc1 epilogue, GELU' fused into the
    fc2 input-gradient epilogue."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, mode):
        comp, cdt = _cfg(mode)
        shp = x.shape
        x2 = K.cast(x.reshape(-1, shp[-1]), cdt) if x.dtype != cdt else x.reshape(-1, shp[-1])
        R, Cin = x2.shape
        Hd, Co = w1.shape[0], w2.shape[0]
        g, h = K.gemm(comp, x2, MAJOR_K, _operand(w1, cdt), MAJOR_K, R, Hd, Cin, cdt, epilogue=EPI_GELU_GRAD,
                      bias=_f32(b1), want_aux_dtype=cdt)        # h = GELU'(pre-activation), for the backward
        y, _ = K.gemm(comp, g, MAJOR_K, _operand(w2, cdt), MAJOR_K, R, Co, Hd, cdt, bias=_f32(b2))
        ctx.save_for_backward(x2, w1, w2, h, g)
        ctx.meta = (shp, mode, b1 is not None, b2 is not None)
        return y.view(*shp[:-1], Co)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w1, w2, h, g = ctx.saved_tensors
        shp, mode, hb1, hb2 = ctx.meta
        comp, cdt = _cfg(mode)
        dy2 = dy.reshape(-1, w2.shape[0])
        dy2 = K.cast(dy2, cdt) if dy2.dtype != cdt else dy2
        ni = ctx.needs_input_grad
        dh, dw2, db2 = _linear_backward(comp, cdt, dy2, g, w2, True, ni[3], hb2 and ni[4], dgelu_aux=h)
        dx, dw1, db1 = _linear_backward(comp, cdt, dh, x2, w1, ni[0], ni[1], hb1 and ni[2])
        if dx is not None:
            dx = dx.view(shp)
        return dx, dw1, db1, dw2, db2, None


def mlp(x, w1, b1, w2, b2, mode=None):
    return MlpFunction.apply(x, w1, b1, w2, b2, mode or get_precision())


# ------------------------------------------------------------------------------------------------
class BlockFunction(Function):
    """One DINOv2 transformer block (dinov2/layers/block.py:89-114, eval path :112-113) as a single
    autograd node:  x + ls1(attn(norm1 x));  x + ls2(mlp(norm2 x)).
    Residual stream, LayerNorm statistics and LayerScale in fp32; GEMM/attention operands in the
    compute dtype.  LayerScale+residual are GEMM epilogues; the residual-branch gradient add is
    fused into the LayerNorm backward."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, g1, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b, g2,
                num_heads, eps, mode, grad_on=True):
        comp, cdt = _cfg(mode)
        B, T, C = x.shape
        R = B * T
        dev = x.device
        x2 = x.reshape(R, C)
        x2 = x2 if x2.dtype == torch.float32 else x2.float()
        hd = C // num_heads
        # forward-only calls (the taps pass runs under no_grad: nothing needs a gradient) skip the tensors
        # that exist only for backward: the GELU pre-activation and the LayerScale branch outputs
        # (``grad_on`` is torch.is_grad_enabled() sampled by the caller: inside Function.forward grad mode is
        # always off, and ctx.needs_input_grad reports the parameters' requires_grad even under no_grad())
        track = grad_on and any(ctx.needs_input_grad)
        save_u = [track and g1 is not None and g1.requires_grad, track and g2 is not None and g2.requires_grad]
        gam1 = _f32(g1) if g1 is not None else _ones(C, dev)
        gam2 = _f32(g2) if g2 is not None else _ones(C, dev)

        y1, mean1, rstd1 = K.layernorm_forward(x2, _f32(n1w), _f32(n1b), eps, cdt)
        qkv, _ = K.gemm(comp, y1, MAJOR_K, _operand(qkv_w, cdt), MAJOR_K, R, 3 * C, C, cdt, bias=_f32(qkv_b))
        o, lse = K.attention_forward(comp, qkv.view(B, T, 3 * C), B, T, num_heads, hd)
        x1, u1 = K.gemm(comp, o.view(R, C), MAJOR_K, _operand(proj_w, cdt), MAJOR_K, R, C, C, torch.float32,
                        epilogue=EPI_SCALE_RESIDUAL, bias=_f32(proj_b), gamma=gam1, residual=x2,
                        want_aux_dtype=cdt if save_u[0] else None)
        y2, mean2, rstd2 = K.layernorm_forward(x1, _f32(n2w), _f32(n2b), eps, cdt)
        Hd = fc1_w.shape[0]
        # with a backward to come the epilogue also saves GELU'(pre-activation) (`h`): it shares the erfc / exp
        # evaluation with the activation, and fc2's input-gradient GEMM then only multiplies by it
        g, h = K.gemm(comp, y2, MAJOR_K, _operand(fc1_w, cdt), MAJOR_K, R, Hd, C, cdt,
                      epilogue=EPI_GELU_GRAD if track else EPI_GELU, bias=_f32(fc1_b), want_aux_dtype=cdt if track else None)
        out, u2 = K.gemm(comp, g, MAJOR_K, _operand(fc2_w, cdt), MAJOR_K, R, C, Hd, torch.float32,
                         epilogue=EPI_SCALE_RESIDUAL, bias=_f32(fc2_b), gamma=gam2, residual=x1,
                         want_aux_dtype=cdt if save_u[1] else None)
        ctx.save_for_backward(x2, mean1, rstd1, y1, qkv, o, lse, u1, x1, mean2, rstd2, y2, h, g, u2,
                              n1w, qkv_w, proj_w, g1, n2w, fc1_w, fc2_w, g2)
        ctx.meta = (B, T, C, num_heads, hd, mode, x.dtype,
                    qkv_b is not None, proj_b is not None, fc1_b is not None, fc2_b is not None)
        return out.view(B, T, C)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (x2, mean1, rstd1, y1, qkv, o, lse, u1, x1, mean2, rstd2, y2, h, g, u2,
         n1w, qkv_w, proj_w, g1, n2w, fc1_w, fc2_w, g2) = ctx.saved_tensors
        B, T, C, H, hd, mode, xdt, hb_qkv, hb_proj, hb_fc1, hb_fc2 = ctx.meta
        comp, cdt = _cfg(mode)
        R = B * T
        ni = ctx.needs_input_grad
        dev = dout.device
        d2 = dout.reshape(R, C)
        d2 = d2 if d2.dtype == torch.float32 else d2.float()

        # ---- MLP branch
        if g2 is not None:      # LayerScale backward: du, d gamma and the fc2 bias gradient in one pass over d2
            du2, dg2, dfc2_b = K.layerscale_backward(d2, u2, _f32(g2), cdt, ni[14] and u2 is not None, hb_fc2 and ni[13])
            dh, dfc2_w, _ = _linear_backward(comp, cdt, du2, g, fc2_w, True, ni[12], False, dgelu_aux=h)
        else:
            dg2 = None
            du2 = K.cast(d2, cdt)
            dh, dfc2_w, dfc2_b = _linear_backward(comp, cdt, du2, g, fc2_w, True, ni[12], hb_fc2 and ni[13], dgelu_aux=h)
        dy2, dfc1_w, dfc1_b = _linear_backward(comp, cdt, dh, y2, fc1_w, True, ni[10], hb_fc1 and ni[11])
        dx1, dn2w, dn2b = K.layernorm_backward(dy2, x1, _f32(n2w), mean2, rstd2, d2, ni[8] or ni[9])
        # ---- attention branch
        if g1 is not None:
            du1, dg1, dproj_b = K.layerscale_backward(dx1, u1, _f32(g1), cdt, ni[7] and u1 is not None, hb_proj and ni[6])
            do, dproj_w, _ = _linear_backward(comp, cdt, du1, o.view(R, C), proj_w, True, ni[5], False)
        else:
            dg1 = None
            du1 = K.cast(dx1, cdt)
            do, dproj_w, dproj_b = _linear_backward(comp, cdt, du1, o.view(R, C), proj_w, True, ni[5], hb_proj and ni[6])
        dqkv = K.attention_backward(comp, qkv.view(B, T, 3 * C), o, lse, do.view(B, T, C), B, T, H, hd).view(R, 3 * C)
        dy1, dqkv_w, dqkv_b = _linear_backward(comp, cdt, dqkv, y1, qkv_w, True, ni[3], hb_qkv and ni[4])
        dx, dn1w, dn1b = K.layernorm_backward(dy1, x2, _f32(n1w), mean1, rstd1, dx1, ni[1] or ni[2])
        dx = dx.view(B, T, C)
        if xdt != torch.float32:
            dx = dx.to(xdt)
        return (dx if ni[0] else None, dn1w, dn1b, dqkv_w, dqkv_b, dproj_w, dproj_b, dg1, dn2w, dn2b,
                dfc1_w, dfc1_b, dfc2_w, dfc2_b, dg2, None, None, None, None)


# ------------------------------------------------------------------------------------------------
class PatchEmbedFunction(Function):
    """Conv2d(k = stride = patch) as patch gather + GEMM (patch_embed.py:65-81)."""

    @staticmethod
    def forward(ctx, img, weight, bias, patch, mode):
        comp, cdt = _cfg(mode)
        B, Cin, H, W = img.shape
        Co = weight.shape[0]
        Kd = Cin * patch * patch
        ldk = (Kd + 7) // 8 * 8
        cols = K.patchify(img, patch, cdt, ldk)
        w2 = weight.detach().reshape(Co, Kd)
        if ldk != Kd:
            w2 = torch.nn.functional.pad(w2, (0, ldk - Kd))
        w2 = K.cast(w2.contiguous(), cdt)
        R = cols.shape[0]
        y, _ = K.gemm(comp, cols, MAJOR_K, w2, MAJOR_K, R, Co, ldk, torch.float32, bias=_f32(bias))
        ctx.save_for_backward(cols)
        ctx.meta = (weight.shape, Kd, ldk, mode, bias is not None)
        return y.view(B, (H // patch) * (W // patch), Co)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (cols,) = ctx.saved_tensors
        wshape, Kd, ldk, mode, has_bias = ctx.meta
        comp, cdt = _cfg(mode)
        Co = wshape[0]
        dy2 = K.cast(dy.reshape(-1, Co), cdt)
        dw = db = None
        if ctx.needs_input_grad[1]:
            dwp, _ = K.gemm(comp, dy2, MAJOR_MN, cols, MAJOR_MN, Co, ldk, dy2.shape[0], torch.float32)
            dw = dwp[:, :Kd].reshape(wshape)
        if has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(dy2)
        return None, dw, db, None, None


class DWConvFunction(Function):
    """Token-major depth-wise 3x3 conv over back-to-back maps (+ fused exact GELU)
    (adapter_blocks.py:62-80 followed by ConvFFN.act :96)."""

    @staticmethod
    def forward(ctx, x, weight, bias, maps, fuse_gelu):
        y, pre = K.dwconv3x3_forward(x, weight, bias, maps, fuse_gelu, True)
        ctx.save_for_backward(x, weight, pre)
        ctx.meta = (maps, fuse_gelu)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, weight, pre = ctx.saved_tensors
        maps, fuse_gelu = ctx.meta
        dx, dw, db = K.dwconv3x3_backward(dy, pre, x, weight, maps, fuse_gelu)
        return dx, dw.view(weight.shape).to(weight.dtype), db, None, None


class Upsample2xFunction(Function):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) of FeatureDecoder
    (backbones/decoders.py:104-127) on an NCHW tensor, computed channels-last by the CUDA kernels."""

    @staticmethod
    def forward(ctx, x):
        xh = x.permute(0, 2, 3, 1).contiguous()          # no copy when x is channels_last
        return K.upsample2x_forward(xh).permute(0, 3, 1, 2)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        gh = gy.permute(0, 2, 3, 1).contiguous()
        return K.upsample2x_backward(gh).permute(0, 3, 1, 2)


def upsample2x(x):
    return Upsample2xFunction.apply(x)


class Upsample2xNHWCFunction(Function):
    """the same resize on a channels-last map [B, H, W, C] -> [B, 2H, 2W, C] (adaptersis_b200/decoders.py)."""

    @staticmethod
    def forward(ctx, x, pad_in=0, pad_out=0):
        ctx.pads = (pad_in, pad_out)
        return K.upsample2x_forward(x.contiguous(), pad_in, pad_out)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        # (gy has the forward's dtype: produced by our own nodes; only its logical pixels are read)
        return K.upsample2x_backward(gy.contiguous(), *ctx.pads), None, None


def upsample2x_nhwc(x, pad_in=0, pad_out=0):
    """storage padding: x with pad_in border pixels, the result with pad_out zero border pixels (feeds the implicit GEMM)."""
    return Upsample2xNHWCFunction.apply(x, pad_in, pad_out)
