"""Run one kernel a few times (for `ncu --set full -k regex:...`)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptersis_b200 import kernels as K  # noqa: E402
from adaptersis_b200._lib import BF16, EPI_GELU, EPI_GELU_GRAD, EPI_NONE, EPI_SCALE_RESIDUAL, MAJOR_K, MAJOR_MN  # noqa: E402

dev = "cuda"
which = sys.argv[1]
R = 21180
torch.manual_seed(0)
if which.startswith("gemm"):
    A = torch.randn(R, 1024, device=dev).bfloat16()
    W = torch.randn(4096, 1024, device=dev).bfloat16()
    bias = torch.randn(4096, device=dev)
    for _ in range(3):
        if which == "gemm_gelu":
            K.gemm(BF16, A, MAJOR_K, W, MAJOR_K, R, 4096, 1024, torch.bfloat16, epilogue=EPI_GELU, bias=bias, want_aux_dtype=torch.bfloat16)
        elif which == "gemm_gelu_grad":      # fc1 of the training pass: GELU out + its derivative saved (the step's top shape)
            K.gemm(BF16, A, MAJOR_K, W, MAJOR_K, R, 4096, 1024, torch.bfloat16, epilogue=EPI_GELU_GRAD, bias=bias, want_aux_dtype=torch.bfloat16)
        elif which == "gemm_proj":           # attention projection: K = N = 1024, LayerScale + fp32 residual epilogue, saved branch output
            W2 = W[:1024].contiguous()
            res = torch.randn(R, 1024, device=dev)
            gam = torch.randn(1024, device=dev)
            K.gemm(BF16, A, MAJOR_K, W2, MAJOR_K, R, 1024, 1024, torch.float32, epilogue=EPI_SCALE_RESIDUAL, bias=bias[:1024].contiguous(),
                   gamma=gam, residual=res, want_aux_dtype=torch.bfloat16)
        else:
            K.gemm(BF16, A, MAJOR_K, W, MAJOR_K, R, 4096, 1024, torch.bfloat16)
elif which.startswith("attn"):
    B, T, H = 12, 1765, 16
    qkv = torch.randn(B, T, 3 * H * 64, device=dev).bfloat16()
    for _ in range(3):
        out, lse = K.attention_forward(BF16, qkv, B, T, H, 64)
    if which == "attn_bwd":
        dout = torch.randn_like(out)
        for _ in range(3):
            K.attention_backward(BF16, qkv, out, lse, dout, B, T, H, 64)
elif which == "conv":                        # stem 64 -> 64 at 294 x 294, the implicit GEMM (forward, dgrad, wgrad)
    Bc, Hc, Cc = 12, 294, 64
    xp = torch.zeros(Bc, Hc + 2, Hc + 2, Cc, device=dev, dtype=torch.bfloat16)
    xp[:, 1:-1, 1:-1] = torch.randn(Bc, Hc, Hc, Cc, device=dev).bfloat16()
    w2 = torch.randn(Cc, 9 * Cc, device=dev).bfloat16()
    yp = torch.empty_like(xp)
    dw = torch.empty(Cc, 9 * Cc, device=dev)
    for _ in range(2):
        K.conv3x3s1_gemm(0, xp, w2, yp, None, Bc, Hc, Hc, Cc, Cc)
        K.conv3x3s1_gemm(1, xp, w2, yp, None, Bc, Hc, Hc, Cc, Cc)
        K.conv3x3s1_gemm(2, xp, xp, dw, None, Bc, Hc, Hc, Cc, Cc)
elif which == "msda":                        # injector shape, bf16 values (the form the step runs): forward + backward
    from tools.msda_bench import make
    shapes = [(73, 73), (36, 36), (18, 18)]
    value, ss, lsi, loc, aw, gout = make(12, 1764, 8, 128, shapes, 4, torch.bfloat16, dev)
    for _ in range(2):
        K.msda_forward(value, ss, lsi, loc, aw)
        K.msda_backward(value, ss, lsi, loc, aw, gout)
elif which == "ln_fwd":
    x = torch.randn(R, 1024, device=dev)
    w = torch.randn(1024, device=dev)
    b = torch.randn(1024, device=dev)
    for _ in range(3):
        K.layernorm_forward(x, w, b, 1e-6, torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        K.layernorm_forward(x, w, b, 1e-6, torch.bfloat16)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    print(f"ln_fwd {t * 1e3:.1f} us  {R * 1024 * 6 / t / 1e6:.0f} GB/s")
elif which == "ln_bwd":
    x = torch.randn(R, 1024, device=dev)
    dy = torch.randn(R, 1024, device=dev).bfloat16()
    dres = torch.randn(R, 1024, device=dev)
    w = torch.randn(1024, device=dev)
    b = torch.randn(1024, device=dev)
    y, mean, rstd = K.layernorm_forward(x, w, b, 1e-6, torch.bfloat16)
    for _ in range(3):
        K.layernorm_backward(dy, x, w, mean, rstd, dres)
    import time
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        K.layernorm_backward(dy, x, w, mean, rstd, dres)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    nb = R * 1024 * (2 + 4 + 4 + 4)
    print(f"ln_bwd {t * 1e3:.1f} us  {nb / t / 1e6:.0f} GB/s")
torch.cuda.synchronize()
print("done", which)
