"""Run one kernel a few times (for `ncu --set full -k regex:...`)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptersis_b200 import kernels as K  # noqa: E402
from adaptersis_b200._lib import BF16, EPI_GELU, EPI_NONE, MAJOR_K, MAJOR_MN  # noqa: E402

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
