"""Per-tile event timeline of the attention backward dK/dV kernel (CTA 0,0,0), from clock64 stamps.
    ASIS_TRACE=1 python -m adaptersis_b200.build;  ASIS_LIB=adaptersis_b200/libasis_b200_trace.so python tools/attn_bwd_trace.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptersis_b200 import _lib, kernels as K  # noqa: E402
from adaptersis_b200._lib import BF16  # noqa: E402

lib = ctypes.CDLL(_lib.LIB_PATH)
B, T, H = 12, 1765, 16
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").bfloat16()
out, lse = K.attention_forward(BF16, qkv, B, T, H, 64)
dout = torch.randn_like(out)
for _ in range(2):
    K.attention_backward(BF16, qkv, out, lse, dout, B, T, H, 64)
buf = torch.zeros(16, 1024, dtype=torch.int64, device="cuda")
assert lib.asis_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr())) == 0
K.attention_backward(BF16, qkv, out, lse, dout, B, T, H, 64)
torch.cuda.synchronize()
assert lib.asis_debug_set_attn_trace(ctypes.c_void_p(0)) == 0
t = buf.cpu()
names = {0: "mma:P ready", 1: "mma:G2 issued", 2: "mma:S issued", 4: "sm:begin", 5: "sm:S ready", 6: "sm:ld done", 9: "sm:computed",
         10: "sm:arrived"}
t0 = int(t[t > 0].min())
print("tile " + " ".join(f"{names[s]:>13s}" for s in sorted(names)))
for g in range(28):
    print(f"{g:4d} " + " ".join(f"{(int(t[s, g]) - t0) if t[s, g] > 0 else -1:13d}" for s in sorted(names)))
