"""Stress the bf16 attention kernels: repeat on the same input, compare bitwise with the first
result, look for non-finite values; interleave with GEMMs to vary what runs before/after."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptersis_b200 import kernels as K  # noqa: E402
from adaptersis_b200._lib import BF16, MAJOR_K  # noqa: E402

dev = "cuda"
B, T, H = 12, 1765, 16
torch.manual_seed(0)
qkv = (torch.randn(B, T, 3 * H * 64, device=dev) * 0.5).bfloat16()
A = torch.randn(B * T, 1024, device=dev).bfloat16()
W = torch.randn(4096, 1024, device=dev).bfloat16()
out0, lse0 = K.attention_forward(BF16, qkv, B, T, H, 64)
dout = torch.randn_like(out0)
g0 = K.attention_backward(BF16, qkv, out0, lse0, dout, B, T, H, 64)
torch.cuda.synchronize()
print("first finite:", bool(torch.isfinite(out0.float()).all()), bool(torch.isfinite(lse0).all()), bool(torch.isfinite(g0.float()).all()))
bad_f = bad_b = 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for i in range(n):
    if i % 3 == 0:
        K.gemm(BF16, A, MAJOR_K, W, MAJOR_K, B * T, 4096, 1024, torch.bfloat16)
    out, lse = K.attention_forward(BF16, qkv, B, T, H, 64)
    if i % 2 == 0:
        g = K.attention_backward(BF16, qkv, out, lse, dout, B, T, H, 64)
        if not torch.equal(g, g0):
            bad_b += 1
    if not (torch.equal(out, out0) and torch.equal(lse, lse0)):
        bad_f += 1
        if bad_f <= 3:
            d = (out.float() - out0.float()).abs()
            print(f"iter {i}: fwd mismatch max {float(d.max()):.3e} nonfinite {int((~torch.isfinite(out.float())).sum())} "
                  f"rows {int((d.amax(-1) > 0).sum())}", flush=True)
torch.cuda.synchronize()
print(f"fwd mismatches {bad_f}/{n}  bwd mismatches {bad_b}/{(n + 1) // 2}")
