"""Stress the bf16 attention kernels: repeat the forward on the same input (interleaved with GEMMs and
backward launches), checksum every result, and report how many distinct outputs appeared -- 1 means
bit-identical run to run."""
import collections
import os
import sys
import time
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
wts = torch.randn(B, T, H * 64, device=dev)
sums = collections.Counter()
outs = {}
t0 = time.time()
for i in range(n):
    if i % 3 == 0:
        K.gemm(BF16, A, MAJOR_K, W, MAJOR_K, B * T, 4096, 1024, torch.bfloat16)
    out, lse = K.attention_forward(BF16, qkv, B, T, H, 64)
    if i % 4 == 0:
        K.attention_backward(BF16, qkv, out, lse, out, B, T, H, 64)
    key = (float((out.float() * wts).sum()), float(lse.sum()))
    sums[key] += 1
    if key not in outs:
        outs[key] = out.clone()
torch.cuda.synchronize()
mode = sums.most_common(1)[0][0]
print(f"{n} forwards in {time.time() - t0:.1f} s: {len(sums)} distinct outputs; most common seen {sums[mode]} times; "
      f"finite: {bool(torch.isfinite(outs[mode].float()).all())}")
for k, o in list(outs.items())[:4]:
    if k != mode:
        d = (o.float() - outs[mode].float()).abs()
        print(f"   variant seen {sums[k]}x: max diff {float(d.max()):.3e}, rows differing {int((d.amax(-1) > 0).sum())}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    K.attention_forward(BF16, qkv, B, T, H, 64)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 50 * 1e3
print(f"forward {us:.1f} us = {4.0 * B * H * T * T * 64 / us / 1e6:.0f} TFLOP/s (back to back, warm L2)")
