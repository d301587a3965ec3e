"""FeatureDecoder (backbones/decoders.py:92-164) -- same layer layout, constructor arguments and state_dict keys,
running on libasis_b200 kernels (adaptersis_b200/conv.py), channels-last:

  decoder_1..4   3x3 conv (pad 1, bias) + BatchNorm2d + ReLU + bilinear x2 (align_corners=True)    42 -> 84 -> 168 -> 336 -> 672
  final_out      3x3 conv 64 -> num_classes

The four wide convolutions are implicit tcgen05 GEMMs over zero-padded maps in bf16 mode (asis_conv3x3s1_gemm; the
first one, 3 x embed_dim -> 512 at 42 x 42, is a K = 27648 GEMM), asis_im2col + FFMA GEMM in fp32 parity mode; BatchNorm + ReLU is one statistics pass + one fused normalise pass; the resize is the channels-last
kernel pair of csrc/misc.cu; the last resize and the head are one fused node (conv.SegHeadFunction).  Plain BatchNorm2d: per-GPU statistics under data
parallelism, as in the reference."""
import torch
import torch.nn as nn

from . import conv as Cv
from . import functional as Fn


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

    def forward(self, x, record=None):
        """x [B, 3*embed_dim, h, w] (any memory format; channels-last is free) -> logits [B, num_classes, 16h, 16w] f32.
        ``record`` (parity tests): dict that receives every layer's ReLU mask under its BatchNorm's key prefix."""
        x = x.permute(0, 2, 3, 1).contiguous()          # (asis_im2col converts to the compute dtype on the way)
        fused_head = self.num_classes <= 4
        imp = [int(Cv.implicit_ok(getattr(self, f"decoder_{k}")[0])) for k in range(1, 5)] + [0]
        if imp[0]:          # bf16 mode: zero-padded bf16 storage, the convolutions are implicit GEMMs (no column matrix)
            x = Cv.repad(x, 0, 1, torch.bfloat16)
        for k in range(1, 5):
            conv, bn, _, _ = getattr(self, f"decoder_{k}")
            x = Cv.conv_bn_relu(x, conv, bn, imp[k - 1], 0, record, f"decoder_{k}.1.")
            if k < 4 or not fused_head:
                x = Fn.upsample2x_nhwc(x, 0, imp[k] if k < 4 else 0)
        if fused_head:      # decoder_4's resize + final_out as one node: the 64-channel 672 x 672 map is never built
            y = Cv.seg_head(x, self.final_out.weight, self.final_out.bias)
        else:
            y = Cv.conv2d(x, self.final_out.weight, self.final_out.bias, 1, 1, out_dtype=torch.float32)
        return y.permute(0, 3, 1, 2)
