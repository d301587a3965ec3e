"""Debug: run a few train steps and report the loss and any non-finite gradient per step."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from adaptersis_b200.trainer import TrainStep  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.manual_seed(0)
ts = TrainStep(arch="vit_large", device="cuda", precision="bf16")
batches = [bench.synth_batch(B, 588, 2, i) for i in range(2)]
for i in range(steps):
    inp, tgt = (t.cuda() for t in batches[i % 2])
    ts.optimizer.zero_grad(set_to_none=True)
    loss = ts.forward_loss(inp, tgt)
    loss.backward()
    torch.cuda.synchronize()
    bad = [(n, float(p.grad.abs().max())) for n, p in ts.named_parameters()
           if p.grad is not None and not torch.isfinite(p.grad).all()]
    big = sorted(((float(p.grad.abs().max()), n) for n, p in ts.named_parameters() if p.grad is not None), reverse=True)[:3]
    print(f"step {i} loss {float(loss):.6f} nonfinite grads: {len(bad)} {bad[:4]} largest {big}", flush=True)
    ts.reducer.finish()
    ts.optimizer.step()
