"""One line per call: forward / backward times of the two real MSDeformAttn shapes (bf16 and fp32) with the library
ASIS_LIB names (tools/msda_variants.sh), plus a bit-exactness check of the outputs against a checksum printed beside
them (the variants must not change a single bit: same summation order)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptersis_b200 import kernels as K  # noqa: E402
import msda_bench as mb  # noqa: E402


def main(tag):
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"variant": tag}
    for name, N, Lq, M, D, shapes, P in mb.REAL_CASES:
        for dtype in (torch.bfloat16, torch.float32):
            qs = [(73, 73), (36, 36), (18, 18)] if name == "extractor_real" else [(42, 42)]
            v, ss, lsi, loc, aw, gout = mb.make(N, Lq, M, D, shapes, P, dtype, dev, qgrids=qs)
            tf = mb.timeit(lambda: K.msda_forward(v, ss, lsi, loc, aw), 15, flush)
            tb = mb.timeit(lambda: K.msda_backward(v, ss, lsi, loc, aw, gout), 15, flush)
            o = K.msda_forward(v, ss, lsi, loc, aw)
            gv, gl, ga = K.msda_backward(v, ss, lsi, loc, aw, gout)
            chk = [int(t.contiguous().view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32).long().sum()) & 0xffffffff
                   for t in (o, gv, gl, ga)]
            key = f"{name[:3]}_{'bf16' if dtype == torch.bfloat16 else 'f32'}"
            out[key] = {"fwd_us": round(tf * 1e3, 1), "bwd_us": round(tb * 1e3, 1), "chk": chk}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "base")
