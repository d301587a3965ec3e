"""Kernel-time breakdown of one bench step with torch.profiler (CUPTI) -- quick look at what is
NOT one of our kernels.  (The committed launch lists come from ncu; this is a development aid.)"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from adaptersis_b200.trainer import TrainStep  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
ts = TrainStep(arch="vit_large", device=dev, precision="bf16")
a, b = bench.synth_batch(12, 588, 2, 0)
a, b = a.to(dev), b.to(dev)
for _ in range(3):
    ts.step_device(a, b)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ts.step_device(a, b)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t:
        rows.append((t / 1e3, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total kernel time {tot:.1f} ms")
ours = sum(r[0] for r in rows if "asis::" in r[2])
print(f"asis:: kernels {ours:.1f} ms ({100 * ours / tot:.1f}%)")
for t, n, k in rows[:60]:
    print(f"{t:8.2f} ms {100 * t / tot:5.1f}% n={n:4d} {k[:110]}")
