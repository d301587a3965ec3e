"""The interleaved ViT + adapter encoder that the reference writes as straight-line code inside
train() / validate_network() (train.py:275-406), as one nn.Module.

  taps  = feature_model(img)            -> 4 last blocks, CLS + pos-embed, final norm, no grad (F6)
  x     = patch_embed(img) -> blocks[0:-3]                 (no CLS, no pos-embed: train.py:300-302)
  4 x : [blocks[-3+k]] -> injector(x, c) -> extractor(c, x) -> x + tap_k
  feat  = cat(out_last, pad(c4), out_vit)  [B, 3C, 42, 42]

``frozen_backbone=True`` reproduces the reference wiring (backbone blocks under no_grad, graph cut
before the decoder); ``False`` (default) keeps the graph connected so the backbone
runs forward *and* backward, which is what BASELINE.json measures."""
import contextlib

import torch
import torch.nn as nn
import torch.nn.functional as F

from .adapter_blocks import CACNN, CAViT, deform_inputs
from .encoders import FeatureEncoder
from .vision_transformer import ModelWithIntermediateLayers, build_model_for_eval


class AdapterEncoder(nn.Module):
    def __init__(self, arch="vit_large", embed_dim=None, adapter_heads=8, n_points=4, inplanes=64, patch_size=14,
                 n_last_blocks=4, frozen_backbone=False, model=None, injector_init=0.0):
        super().__init__()
        self.model = model if model is not None else build_model_for_eval(arch, patch_size=patch_size)
        if self.model.chunked_blocks:
            raise ValueError("AdapterEncoder indexes model.blocks[i] like train.py:300-370: build the backbone with "
                             "block_chunks=0 (build_model_for_eval does)")
        C = self.model.embed_dim
        self.patch_size = self.model.patch_size
        self.feature_model = ModelWithIntermediateLayers(self.model, n_last_blocks)
        self.backbone_encoder = FeatureEncoder(inplanes=inplanes, embed_dim=C)
        self.cross_vit = CAViT(dim=C, n_levels=3, num_heads=adapter_heads, n_points=n_points, init_values=injector_init)
        self.cross_cnn = CACNN(dim=C, n_levels=1, num_heads=adapter_heads, n_points=n_points, with_cffn=True,
                               cffn_ratio=0.25)
        self.frozen_backbone = frozen_backbone

    def forward(self, inp):
        B, _, H, W = inp.shape
        model = self.model
        d1, d2 = deform_inputs(inp, self.patch_size)
        H_c, W_c = H // 16, W // 16
        _, c2, c3, c4 = self.backbone_encoder(inp, need_c1=False)
        c = torch.cat([c2, c3, c4], dim=1)
        taps = [t for t, _ in self.feature_model(inp)]
        guard = torch.no_grad if self.frozen_backbone else contextlib.nullcontext
        depth = len(model.blocks)
        with guard():
            x = model.patch_embed(inp)
            for blk in model.blocks[0:depth - 3]:
                x = blk(x)
        for stage in range(4):
            if stage > 0:
                with guard():
                    x = model.blocks[depth - 4 + stage](x)
            # (return_feat: the value-side tensor comes back attached to the module's feat_norm node, so the gradients
            #  of its later uses are added inside that LayerNorm's backward kernel instead of by separate ATen passes)
            x, c = self.cross_vit(query=x, reference_points=d1[0], feat=c, spatial_shapes=d1[1], level_start_index=d1[2],
                                  return_feat=True)
            c, x = self.cross_cnn(query=c, reference_points=d2[0], feat=x, spatial_shapes=d2[1], level_start_index=d2[2],
                                  H=H_c, W=W_c, return_feat=True)
            x = x + taps[stage]
        gh, gw = H // self.patch_size, W // self.patch_size
        C = x.shape[-1]
        with guard():
            # train.py:389-406 -- rearrange "b (h w) c -> b c h w", pad c4 18 -> 42, cat along the channels.  The token
            # tensors ARE the channels-last maps, so the concatenation is built channels-last ([B, gh, gw, 3C]) and
            # handed on as the NCHW tensor the reference produces: same logical values, no transposes; the decoder's
            # permute back to channels-last is a view.
            s4 = H // 32
            dy, dx = gh - s4, gw - s4
            c4m = F.pad(c4.reshape(B, s4, s4, C), [0, 0, dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
            feat = torch.cat((x.reshape(B, gh, gw, C), c4m, taps[3].reshape(B, gh, gw, C)), dim=3).permute(0, 3, 1, 2)
        return {"x": x, "c": c, "feat": feat}
