"""Per-tile event timeline of the attention forward kernel v8 (CTA 0, both warpgroups), from clock64 stamps.
Needs a trace build:  ASIS_TRACE=1 python -m adaptersis_b200.build   (a separate library, libasis_b200_trace.so);
run with ASIS_LIB=adaptersis_b200/libasis_b200_trace.so."""
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
for _ in range(2):
    K.attention_forward(BF16, qkv, B, T, H, 64)
buf = torch.zeros(16, 1024, dtype=torch.int64, device="cuda")
assert lib.asis_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr())) == 0
K.attention_forward(BF16, qkv, B, T, H, 64)
torch.cuda.synchronize()
t = buf.cpu()
names = {0: "mma:P0", 1: "mma:P1", 2: "mma:S", 3: "w0:begin", 4: "w0:S", 5: "w0:ld", 6: "w0:max", 7: "w0:exp", 8: "w0:P",
         9: "w1:begin", 10: "w1:S", 11: "w1:ld", 12: "w1:max", 13: "w1:exp", 14: "w1:P"}
t0 = int(t[t > 0].min())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 44
print("tile " + " ".join(f"{names[s]:>9s}" for s in sorted(names)))
for g in range(n):
    print(f"{g:4d} " + " ".join(f"{(int(t[s, g]) - t0) if t[s, g] > 0 else -1:9d}" for s in sorted(names)))
for w in (0, 1):
    b = 3 + 6 * w
    d = lambda a, c: float((t[c, 2:n] - t[a, 2:n]).float().mean())
    print(f"w{w}: wait S {d(b, b + 1):.0f}  load {d(b + 1, b + 2):.0f}  max {d(b + 2, b + 3):.0f}  exp {d(b + 3, b + 4):.0f}  wait O + store {d(b + 4, b + 5):.0f}"
          f"  period {float((t[b, 3:n] - t[b, 2:n - 1]).float().mean()):.0f}")
