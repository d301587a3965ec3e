"""Spatial prior module ("FeatureEncoder", backbones/encoders.py:4-74) -- same layer layout, constructor
arguments and state_dict keys (nn.Conv2d / nn.SyncBatchNorm hold the parameters and running statistics), running
on libasis_b200 kernels (adaptersis_b200/conv.py, csrc/conv.cu), channels-last from the first convolution on:

  stem   3x3/s2 conv (3 -> p) + SyncBN + ReLU, 2 x [3x3 conv + SyncBN + ReLU], 3x3/s2 max pool       588 -> 294 -> 147
  conv2..4   3x3/s2 conv + SyncBN + ReLU (padding 0, 0, 1: only 588 lines the pyramid up)     73, 36, 18
  fc1..4     1x1 convolutions to embed_dim = one GEMM each on the token-major maps: the reference's
             ``c.view(bs, dim, -1).transpose(1, 2)`` is the identity here (the maps already are [B, H*W, C])

The two stride-1 convolutions of the stem run as implicit GEMMs over zero-padded maps in bf16 mode
(asis_conv3x3s1_gemm), every other convolution is asis_im2col + the tcgen05 GEMM (bf16 mode) or the FFMA GEMM (fp32
parity mode); every
SyncBatchNorm exchanges one small tensor per direction (global-batch statistics, the reference's semantics).
``with_cp`` (activation checkpointing; the reference's own implementation of it is broken, encoders.py:71) is
accepted and ignored."""
import torch
import torch.nn as nn

from . import conv as Cv
from . import functional as Fn


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

    @staticmethod
    def _project(fc, c):
        """1x1 convolution of a channels-last map = nn.Linear over its pixels -> tokens [B, H*W, embed_dim] (f32)."""
        B, H, W, C = c.shape
        return Fn.linear(c.reshape(B, H * W, C), fc.weight.view(fc.out_channels, C), fc.bias, out_dtype=torch.float32)

    def forward(self, x, need_c1=True, record=None):
        """x [B, 3, H, W] (as the reference) -> (c1 [B, D, H/4, W/4] or None, c2, c3, c4 tokens [B, n, D]).
        ``record`` (parity tests): dict that receives every layer's ReLU mask under its BatchNorm's key prefix."""
        x = x.permute(0, 2, 3, 1).contiguous().float()                   # channels-last image
        s = self.stem
        p3, p6 = int(Cv.implicit_ok(s[3])), int(Cv.implicit_ok(s[6]))     # bf16 mode: implicit GEMMs over zero-padded maps
        x = Cv.conv_bn_relu(x, s[0], s[1], 0, p3, record, "stem.1.")
        x = Cv.conv_bn_relu(x, s[3], s[4], p3, p6, record, "stem.4.")
        x = Cv.conv_bn_relu(x, s[6], s[7], p6, 0, record, "stem.7.")
        c1 = Cv.maxpool3x3s2(x)
        c2 = Cv.conv_bn_relu(c1, self.conv2[0], self.conv2[1], 0, 0, record, "conv2.1.")
        c3 = Cv.conv_bn_relu(c2, self.conv3[0], self.conv3[1], 0, 0, record, "conv3.1.")
        c4 = Cv.conv_bn_relu(c3, self.conv4[0], self.conv4[1], 0, 0, record, "conv4.1.")
        o1 = None
        if need_c1:        # fc1(c1) feeds nothing downstream in train.py (:279, then unused); need_c1=False skips it
            B, H, W, _ = c1.shape
            o1 = self._project(self.fc1, c1).view(B, H, W, -1).permute(0, 3, 1, 2)
        return o1, self._project(self.fc2, c2), self._project(self.fc3, c3), self._project(self.fc4, c4)
