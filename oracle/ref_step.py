"""The reference's OWN modules (staged unmodified under oracle/_ref/ by oracle/make_ref.sh) composed into
one training iteration on the host CPU -- TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's `--impl reference`
arm and `cpu_baseline`; never imported by the product package).

The composition is the data flow of train.py:275-436 (F6: taps pass through `get_intermediate_layers` under
no_grad, then patch_embed -> blocks[0:-3] -> 4 x (block, CAViT, CACNN, + tap) -> concat -> FeatureDecoder ->
bilinear resize -> Softmax -> DC dice loss -> backward), with the same three deviations from "as shipped" that
tests/golden/make_golden.py documents, all forced by running on a CPU with gradients:
  * the graph is left connected (no torch.no_grad around the backbone blocks / the concat): BASELINE.json
    measures backbone forward AND backward;
  * `MSDeformAttnFunction` has no backward (F2): its `apply` is pointed at the reference's own
    `ms_deform_attn_core_pytorch` (plain autograd through F.grid_sample);
  * `nn.SyncBatchNorm` cannot run on CPU: converted to `nn.BatchNorm2d` with identical parameters;
  * `DC.onehot` calls `.cuda()`: the loss is handed an already one-hot target, which skips that call
    (segloss/dice.py:23-25) -- the arithmetic is DC.dice itself.
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "backbones", "adapter_blocks.py"))


def _import_reference():
    os.environ.setdefault("XFORMERS_DISABLED", "1")       # the reference's own switch: naive attention path
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from backbones.ops.modules import ms_deform_attn as ref_msda
        from backbones import adapter_blocks as ref_ab
        from backbones.encoders import FeatureEncoder
        from backbones.decoders import FeatureDecoder
        from dinov2.layers import MemEffAttention, NestedTensorBlock
        from dinov2.models import vision_transformer as ref_vits
        from segloss.dice import DC
    return dict(msda=ref_msda, ab=ref_ab, FeatureEncoder=FeatureEncoder, FeatureDecoder=FeatureDecoder,
                MemEffAttention=MemEffAttention, Block=NestedTensorBlock, vits=ref_vits, DC=DC)


def _sync_bn_to_bn(module):
    for name, child in module.named_children():
        if isinstance(child, nn.SyncBatchNorm):
            bn = nn.BatchNorm2d(child.num_features, eps=child.eps, momentum=child.momentum)
            bn.load_state_dict(child.state_dict())
            setattr(module, name, bn)
        else:
            _sync_bn_to_bn(child)
    return module


ARCH = {"vit_small": "vit_small", "vit_base": "vit_base", "vit_large": "vit_large"}


class ReferenceStep:
    """Reference modules at the benchmark's dimensions, random init (the reference constructors' own)."""

    def __init__(self, arch="vit_large", adapter_heads=8, num_classes=2, seed=0):
        from functools import partial
        R = _import_reference()
        self.R = R
        torch.manual_seed(seed)
        # what dinov2/eval/setup.py:62-67 + dinov2/models/__init__.py:14-40 build for configs/eval/vit*14_pretrain.yaml
        self.model = getattr(R["vits"], ARCH[arch])(patch_size=14, img_size=518, init_values=1e-5, ffn_layer="mlp",
                                                    block_chunks=0, qkv_bias=True, proj_bias=True, ffn_bias=True)
        self.model.eval()
        C = self.model.embed_dim
        self.C = C
        self.spm = _sync_bn_to_bn(R["FeatureEncoder"](inplanes=64, embed_dim=C))
        self.inj = R["ab"].CAViT(dim=C, n_levels=3, num_heads=adapter_heads, n_points=4, init_values=0.0)
        self.ext = R["ab"].CACNN(dim=C, n_levels=1, num_heads=adapter_heads, n_points=4, with_cffn=True, cffn_ratio=0.25)
        self.dec = R["FeatureDecoder"](embed_dim=C, num_classes=num_classes, features=[C, 512, 256, 128, 64])
        self.loss = R["DC"](num_classes)
        self.num_classes = num_classes
        self.modules = [self.model, self.spm, self.inj, self.ext, self.dec]

        core = R["msda"].ms_deform_attn_core_pytorch

        class _CoreApply:                                   # autograd-capable stand-in (see module docstring)
            @staticmethod
            def apply(value, shapes, lsi, loc, aw, im2col_step):
                return core(value, shapes, loc, aw)
        self._core_apply = _CoreApply

    def step(self, img, target):
        """forward + backward of one batch; returns the loss as a python float."""
        R = self.R
        for m in self.modules:
            for p in m.parameters():
                p.grad = None
        model, depth = self.model, len(self.model.blocks)
        old = R["msda"].MSDeformAttnFunction
        R["msda"].MSDeformAttnFunction = self._core_apply
        try:
            B, _, H, W = img.shape
            d1, d2 = R["ab"].deform_inputs(img, 14)
            c1, c2, c3, c4 = self.spm(img)
            c = torch.cat([c2, c3, c4], dim=1)
            with torch.no_grad():
                taps = [t for t, _ in model.get_intermediate_layers(img, 4, return_class_token=True)]
            x = model.patch_embed(img)
            for blk in model.blocks[0:-3]:
                x = blk(x)
            for stage in range(4):
                if stage > 0:
                    x = model.blocks[depth - 4 + stage](x)
                x = self.inj(query=x, reference_points=d1[0], feat=c, spatial_shapes=d1[1], level_start_index=d1[2])
                c = self.ext(query=c, reference_points=d2[0], feat=x, spatial_shapes=d2[1], level_start_index=d2[2],
                             H=H // 16, W=W // 16)
                x = x + taps[stage]
            g = H // 14
            out_last = x.transpose(1, 2).reshape(B, self.C, g, g)
            out_vit = taps[3].transpose(1, 2).reshape(B, self.C, g, g)
            s4 = H // 32
            c4m = c4.transpose(1, 2).reshape(B, self.C, s4, s4)
            pad = g - s4
            c4m = F.pad(c4m, [pad // 2, pad - pad // 2, pad // 2, pad - pad // 2])
            feat = torch.cat((out_last, c4m, out_vit), dim=1)
            logits = F.interpolate(self.dec(feat), size=(H, W), mode="bilinear")
            prob = nn.Softmax(1)(logits)
            onehot = torch.zeros_like(prob).scatter_(1, target.unsqueeze(1).long(), 1.0)
            loss = self.loss(prob, onehot)
            loss.backward()
        finally:
            R["msda"].MSDeformAttnFunction = old
        return float(loss)
