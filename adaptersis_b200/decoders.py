"""FeatureDecoder (backbones/decoders.py:92-164) -- same layer layout and state_dict keys.
SURVEY.md section 8(f) rank 2: *next*, runs on PyTorch library convolutions for now."""
import torch
import torch.nn as nn

from . import functional as Fn
from .functional import get_precision


class FeatureDecoder(nn.Module):
    def __init__(self, img_size=588, inplanes=64, embed_dim=1024, num_classes=2, features=[1024, 512, 256, 128, 64]):
        super().__init__()
        self.img_size = img_size
        self.features = features
        self.inplanes = inplanes
        self.embed_dim = embed_dim
        self.num_classes = num_classes
        chans = [features[0] * 3] + list(features[1:])
        for k in range(1, 5):
            setattr(self, f"decoder_{k}", nn.Sequential(
                nn.Conv2d(chans[k - 1], chans[k], 3, padding=1), nn.BatchNorm2d(chans[k]), nn.ReLU(inplace=True),
                nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)))
        self.final_out = nn.Conv2d(features[4], num_classes, 3, padding=1)

    def forward(self, x):
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=get_precision() != "fp32"):
            return self._forward(x)

    def _forward(self, x):
        lowp = get_precision() == "bf16" and x.is_cuda
        for k in range(1, 5):
            conv, bn, relu, up = getattr(self, f"decoder_{k}")
            x = relu(bn(conv(x)))
            if lowp and x.shape[1] % 8 == 0:
                # bf16 mode: the 2x bilinear resize is our channels-last kernel (ATen's NHWC kernel runs at
                # ~100 GB/s here and autocast would run it in fp32: 13 ms of a 165 ms step)
                x = Fn.upsample2x(x.to(torch.bfloat16))
            else:
                x = up(x)
        return self.final_out(x)
