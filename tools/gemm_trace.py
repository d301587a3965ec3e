"""Per-tile event timeline of the CTA-pair GEMM kernel (leader CTA of cluster 0), from clock64 stamps.
Needs the trace library:  ASIS_TRACE=1 python -m adaptersis_b200.build   (writes libasis_b200_trace.so), then
    ASIS_LIB=adaptersis_b200/libasis_b200_trace.so python tools/gemm_trace.py [shape-name-substring]
Columns (clocks relative to the first stamp): when the MMA warp starts waiting for a free accumulator / gets it /
sees the first k-block in smem / issues the last k-block; when epilogue warp 0 starts waiting for the accumulator /
gets it / has stored its slab; when the producer issues the first / last k-block of the tile."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptersis_b200 import _lib, kernels as K  # noqa: E402
from adaptersis_b200._lib import BF16, EPI_DGELU, EPI_GELU, EPI_SCALE_RESIDUAL  # noqa: E402
from tools.gemm_bench import SHAPES  # noqa: E402

lib = ctypes.CDLL(_lib.LIB_PATH)
dev = torch.device("cuda:0")
only = sys.argv[1] if len(sys.argv) > 1 else "fc1 fwd +gelu"
names = {7: "mma:wait acc", 0: "mma:acc free", 1: "mma:kb0 ready", 2: "mma:last kb", 3: "epi:wait", 4: "epi:acc full", 8: "epi:done",
         5: "tma:first kb", 6: "tma:last kb"}
order = [7, 0, 1, 2, 3, 4, 8, 5, 6]
def show(t, title, nmax=40):
    t0 = int(t[t > 0].min())
    print(f"== {title}")
    print("tile " + " ".join(f"{names[s]:>13s}" for s in order))
    for g in range(min(nmax, int((t[0] > 0).sum()))):
        print(f"{g:4d} " + " ".join(f"{(int(t[s, g]) - t0) if t[s, g] > 0 else -1:13d}" for s in order))


if only.startswith("conv"):          # conv<op>: the 64 -> 64 implicit convolution of the stem at 294 x 294, 12 images
    op = int(only[4:] or 0)
    Bc, Hc, Cc = 12, 294, 64
    xp = torch.zeros(Bc, Hc + 2, Hc + 2, Cc, device=dev, dtype=torch.bfloat16)
    xp[:, 1:-1, 1:-1] = torch.randn(Bc, Hc, Hc, Cc, device=dev).bfloat16()
    w2 = torch.randn(Cc, 9 * Cc, device=dev).bfloat16()
    yp = torch.empty_like(xp)
    dw = torch.empty(Cc, 9 * Cc, device=dev)
    args = {0: (0, xp, w2, yp), 1: (1, xp, w2, yp), 2: (2, xp, xp, dw)}[op]
    for _ in range(3):
        K.conv3x3s1_gemm(args[0], args[1], args[2], args[3], None, Bc, Hc, Hc, Cc, Cc)
    buf = torch.zeros(9, 256, dtype=torch.int64, device=dev)
    assert lib.asis_debug_set_gemm_trace(ctypes.c_void_p(buf.data_ptr())) == 0
    K.conv3x3s1_gemm(args[0], args[1], args[2], args[3], None, Bc, Hc, Hc, Cc, Cc)
    torch.cuda.synchronize()
    assert lib.asis_debug_set_gemm_trace(ctypes.c_void_p(0)) == 0
    show(buf.cpu(), f"implicit conv op {op}: 12 x 294 x 294, 64 -> 64")
    sys.exit(0)

for name, M, N, Kd, am, bm, epi, odt, aux, bias in SHAPES:
    if only not in name:
        continue
    A = torch.randn((M, Kd) if am == 0 else (Kd, M), device=dev).bfloat16()
    B = torch.randn((N, Kd) if bm == 0 else (Kd, N), device=dev).bfloat16()
    kw = {}
    if bias:
        kw["bias"] = torch.randn(N, device=dev)
    if epi == EPI_SCALE_RESIDUAL:
        kw["gamma"] = torch.randn(N, device=dev)
        kw["residual"] = torch.randn(M, N, device=dev)
    if epi == EPI_DGELU:
        kw["aux"] = torch.randn(M, N, device=dev).bfloat16()
    elif aux:
        kw["want_aux_dtype"] = torch.bfloat16
    out = torch.empty(M, N, dtype=odt, device=dev)
    for _ in range(3):
        K.gemm(BF16, A, am, B, bm, M, N, Kd, odt, epilogue=epi, out=out, **kw)
    buf = torch.zeros(9, 256, dtype=torch.int64, device=dev)
    assert lib.asis_debug_set_gemm_trace(ctypes.c_void_p(buf.data_ptr())) == 0
    K.gemm(BF16, A, am, B, bm, M, N, Kd, odt, epilogue=epi, out=out, **kw)
    torch.cuda.synchronize()
    assert lib.asis_debug_set_gemm_trace(ctypes.c_void_p(0)) == 0
    t = buf.cpu()
    t0 = int(t[t > 0].min())
    print(f"== {name}  M={M} N={N} K={Kd}")
    print("tile " + " ".join(f"{names[s]:>13s}" for s in order))
    for g in range(int((t[0] > 0).sum())):
        print(f"{g:4d} " + " ".join(f"{(int(t[s, g]) - t0) if t[s, g] > 0 else -1:13d}" for s in order))
