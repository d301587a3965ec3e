"""Spatial prior module ("FeatureEncoder", backbones/encoders.py:4-74) -- same layer layout and
state_dict keys.  SURVEY.md section 8(f) rank 1: this component is *next*, not yet on the
hand-written path; it runs on PyTorch library convolutions (cuDNN) + SyncBatchNorm for now and is
excluded from every "our kernels" claim."""
import torch
import torch.nn as nn

from .functional import get_precision


def _cbr(cin, cout, stride, padding):
    return [nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=padding, bias=False),
            nn.SyncBatchNorm(cout), nn.ReLU(inplace=True)]


class FeatureEncoder(nn.Module):
    def __init__(self, inplanes=64, embed_dim=1024, with_cp=False):
        super().__init__()
        self.with_cp = with_cp
        p = inplanes
        self.stem = nn.Sequential(*(_cbr(3, p, 2, 1) + _cbr(p, p, 1, 1) + _cbr(p, p, 1, 1)
                                    + [nn.MaxPool2d(kernel_size=3, stride=2, padding=1)]))
        self.conv2 = nn.Sequential(*_cbr(p, 2 * p, 2, 0))      # padding 0: only 588 lines the pyramid up (F5)
        self.conv3 = nn.Sequential(*_cbr(2 * p, 4 * p, 2, 0))
        self.conv4 = nn.Sequential(*_cbr(4 * p, 8 * p, 2, 1))
        self.fc1 = nn.Conv2d(p, embed_dim, kernel_size=1, bias=True)
        self.fc2 = nn.Conv2d(2 * p, embed_dim, kernel_size=1, bias=True)
        self.fc3 = nn.Conv2d(4 * p, embed_dim, kernel_size=1, bias=True)
        self.fc4 = nn.Conv2d(8 * p, embed_dim, kernel_size=1, bias=True)

    def forward(self, x, need_c1=True):
        # fp32 parity mode: keep the library convolutions out of TF32
        lowp = get_precision() == "bf16" and x.is_cuda
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=get_precision() != "fp32"), \
                torch.autocast("cuda", dtype=torch.bfloat16, enabled=lowp):
            if lowp:
                x = x.contiguous(memory_format=torch.channels_last)
            return self._forward(x, need_c1)

    def _forward(self, x, need_c1):
        c1 = self.stem(x)
        c2 = self.conv2(c1)
        c3 = self.conv3(c2)
        c4 = self.conv4(c3)
        # fc1(c1) feeds nothing downstream in train.py (:279, then unused); need_c1=False skips it
        o1 = self.fc1(c1) if need_c1 else None
        toks = [f(c).flatten(2).transpose(1, 2) for f, c in ((self.fc2, c2), (self.fc3, c3), (self.fc4, c4))]
        return o1, toks[0], toks[1], toks[2]
