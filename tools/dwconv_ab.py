"""Forward / backward time of the token-major depth-wise 3x3 convolution (+ fused GELU) at the step's shape
([12, 6949, 256] bf16 over the 73² + 36² + 18² maps) with the library ASIS_LIB names; checksums of the outputs."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptersis_b200 import functional as Fn  # noqa: E402
import msda_bench as mb  # noqa: E402


def main(tag):
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    maps = [(73, 73), (36, 36), (18, 18)]
    ntok, C = sum(h * w for h, w in maps), 256
    g = torch.Generator().manual_seed(1)
    x = torch.randn(12, ntok, C, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
    w = (torch.randn(C, 1, 3, 3, generator=g) * 0.3).to(dev).requires_grad_(True)
    b = torch.randn(C, generator=g).to(dev).requires_grad_(True)
    dy = torch.randn(12, ntok, C, generator=g).to(dev, torch.bfloat16)
    with Fn.precision("bf16"):
        tf = mb.timeit(lambda: Fn.DWConvFunction.apply(x, w, b, maps, True), 20, flush)
        y = Fn.DWConvFunction.apply(x, w, b, maps, True)
        tb = mb.timeit(lambda: torch.autograd.grad(y, (x, w, b), dy, retain_graph=True), 20, flush)
        gx, gw, gb = torch.autograd.grad(y, (x, w, b), dy, retain_graph=True)
    chk = [int(t.contiguous().view(torch.int16).long().sum()) & 0xffffffff for t in (y, gx)]
    print(json.dumps({"variant": tag, "fwd_us": round(tf * 1e3, 1), "bwd_us": round(tb * 1e3, 1), "chk_y_dx": chk,
                      "dw_abs_sum": float(gw.abs().sum()), "db_abs_sum": float(gb.abs().sum())}), flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "new")
