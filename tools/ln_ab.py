"""One line per call: LayerNorm backward time at the step's two shapes (dy bf16, x f32, fused residual gradient, parameter
gradients) with the library ASIS_LIB names (tools/ln_variants.sh), and the error against torch's fp32 autograd."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptersis_b200 import kernels as K  # noqa: E402
import msda_bench as mb  # noqa: E402


def main(tag):
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"variant": tag}
    for R in (21168, 83388):
        C = 1024
        g = torch.Generator().manual_seed(R)
        x = torch.randn(R, C, generator=g).to(dev)
        dy = torch.randn(R, C, generator=g).to(dev, torch.bfloat16)
        dres = torch.randn(R, C, generator=g).to(dev)
        w = (torch.rand(C, generator=g) + 0.5).to(dev)
        b = torch.zeros(C, device=dev)
        _, mean, rstd = K.layernorm_forward(x, w, b, 1e-6, torch.bfloat16)
        t = mb.timeit(lambda: K.layernorm_backward(dy, x, w, mean, rstd, dres, True), 20, flush)
        dx, dw, db = K.layernorm_backward(dy, x, w, mean, rstd, dres, True)
        xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
        F.layer_norm(xr, (C,), wr, br, 1e-6).backward(dy.float())
        err = [float((a - r).abs().max() / r.abs().max()) for a, r in ((dx, xr.grad + dres), (dw, wr.grad), (db, br.grad))]
        nbytes = R * C * (2 + 4 + 4 + 4)
        out[f"R{R}"] = {"us": round(t * 1e3, 1), "GBs": round(nbytes / t / 1e6), "err": [float(f"{e:.1e}") for e in err]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "base")
