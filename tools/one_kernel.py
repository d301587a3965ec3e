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
torch.cuda.synchronize()
print("done", which)
