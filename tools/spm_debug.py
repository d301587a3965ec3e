"""diagnostic: where do the spatial-prior-module gradients of the CUDA path leave the stock modules' (fp32 mode)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import adaptersis_b200 as asis
from adaptersis_b200 import conv as Cv
from adaptersis_b200.encoders import FeatureEncoder
from test_gpu_conv import _torch_spm
from conftest import relerr

DEV = "cuda"
torch.manual_seed(6)
ref = _torch_spm(64, 128).to(DEV).double()
ours = FeatureEncoder(inplanes=64, embed_dim=128).to(DEV)
ours.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=True)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 292
x = torch.rand(2, 3, size, size, device=DEV)

# reference with retained intermediates
acts_r = {}
h = ref.stem(x.double()); h.retain_grad(); acts_r["c1"] = h
c2 = ref.conv2(h); c2.retain_grad(); acts_r["c2"] = c2
c3 = ref.conv3(c2); c3.retain_grad(); acts_r["c3"] = c3
c4 = ref.conv4(c3); c4.retain_grad(); acts_r["c4"] = c4
outs_r = [f(c).flatten(2).transpose(1, 2) for f, c in ((ref.fc2, c2), (ref.fc3, c3), (ref.fc4, c4))]
gs = [torch.randn(o.shape, device=DEV) for o in outs_r]
torch.autograd.backward(outs_r, [g.double() for g in gs])

with asis.precision("fp32"):
    xs = x.permute(0, 2, 3, 1).contiguous().float()
    s = ours.stem
    a = Cv.conv_bn_relu(xs, s[0], s[1]); a = Cv.conv_bn_relu(a, s[3], s[4]); a = Cv.conv_bn_relu(a, s[6], s[7])
    c1o = Cv.maxpool3x3s2(a); c1o.retain_grad()
    c2o = Cv.conv_bn_relu(c1o, ours.conv2[0], ours.conv2[1]); c2o.retain_grad()
    c3o = Cv.conv_bn_relu(c2o, ours.conv3[0], ours.conv3[1]); c3o.retain_grad()
    c4o = Cv.conv_bn_relu(c3o, ours.conv4[0], ours.conv4[1]); c4o.retain_grad()
    outs = [ours._project(f, c) for f, c in ((ours.fc2, c2o), (ours.fc3, c3o), (ours.fc4, c4o))]
    torch.autograd.backward(outs, gs)
import copy
r32 = copy.deepcopy(ref).float()
t1 = r32.stem(x); t2 = r32.conv2(t1); t3 = r32.conv3(t2); t4 = r32.conv4(t3)
acts_t = dict(c1=t1, c2=t2, c3=t3, c4=t4)
for n, o in (("c1", c1o), ("c2", c2o), ("c3", c3o), ("c4", c4o)):
    r = acts_r[n]
    on = o.permute(0, 3, 1, 2)
    print(n, tuple(o.shape), "act", relerr(on, r), "(stock fp32:", relerr(acts_t[n], r), ") grad", relerr(o.grad.permute(0, 3, 1, 2), r.grad),
          "| ReLU mask flips vs fp64: ours", int(((on > 0) != (r > 0)).sum()), "stock fp32", int(((acts_t[n] > 0) != (r > 0)).sum()), "of", r.numel())
for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
    if p.grad is not None and q.grad is not None:
        print(n, relerr(p.grad, q.grad))
