"""Oracle: spatial prior module, decoder head and the interleaved adapter encoder
(test infrastructure only).

  * ``spm``             <- ``FeatureEncoder.forward`` (backbones/encoders.py:49-74); SyncBatchNorm in
                           training mode == batch statistics over the (global) batch
  * ``feature_decoder`` <- ``FeatureDecoder.forward`` (backbones/decoders.py:92-164), BatchNorm in training mode
  * ``adapter_encoder`` <- the straight-line network inside ``train()`` (train.py:275-406).
      The reference wraps the backbone blocks and the final concat in ``torch.no_grad()``
      (train.py:286,325,346,368,389), so as shipped no gradient reaches adapters or backbone
      (SURVEY.md F3).  BASELINE.json asks for backbone forward *and backward*, therefore this
      oracle composes the *same modules in the same order* without those scopes; only the
      ``feature_model`` taps stay constant (they come out of ``torch.inference_mode`` inside
      ``ModelWithIntermediateLayers``, dinov2/eval/utils.py:38-44).
  * ``dice_loss``       <- ``DC.forward`` (segloss/dice.py:5-36) after ``nn.Softmax(1)`` (train.py:424):
      softmax is applied twice in the reference; restated as such.
"""
import torch
import torch.nn.functional as F

from . import adapter, layers, vit

BN_EPS = 1e-5


def _bn_train(sd, prefix, x):
    mu = x.mean((0, 2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean((0, 2, 3), keepdim=True)
    w = sd[prefix + "weight"].view(1, -1, 1, 1)
    b = sd[prefix + "bias"].view(1, -1, 1, 1)
    return (x - mu) * torch.rsqrt(var + BN_EPS) * w + b


def _cbr(sd, conv, bn, x, stride, padding, masks=None):
    """conv -> BatchNorm (batch statistics) -> ReLU.  ``masks`` (checker option): {bn prefix: bool [B,C,H,W]} replaces
    relu(pre) by pre * mask.  ReLU' is discontinuous at 0: a pre-activation within rounding distance of zero takes the
    other branch in any two fp32 evaluations (cuDNN vs cuDNN with another algorithm included), and one flipped element
    moves a channel's parameter gradients by up to a few per cent on small maps.  A gradient check therefore evaluates
    this oracle ON THE MASK OF THE RUN IT CHECKS (the function is then smooth and 1e-4-comparable); forward parity is
    checked with the oracle's own ReLU."""
    x = F.conv2d(x, sd[conv + "weight"], sd.get(conv + "bias"), stride=stride, padding=padding)
    pre = _bn_train(sd, bn, x)
    if masks is not None and bn in masks:
        return pre * masks[bn].to(pre.dtype)
    return torch.relu(pre)


def spm(sd, img, relu_masks=None):
    m = relu_masks
    x = _cbr(sd, "stem.0.", "stem.1.", img, 2, 1, m)
    x = _cbr(sd, "stem.3.", "stem.4.", x, 1, 1, m)
    x = _cbr(sd, "stem.6.", "stem.7.", x, 1, 1, m)
    c1 = F.max_pool2d(x, 3, 2, 1)
    c2 = _cbr(sd, "conv2.0.", "conv2.1.", c1, 2, 0, m)
    c3 = _cbr(sd, "conv3.0.", "conv3.1.", c2, 2, 0, m)
    c4 = _cbr(sd, "conv4.0.", "conv4.1.", c3, 2, 1, m)
    outs = []
    for i, c in enumerate((c1, c2, c3, c4), 1):
        outs.append(F.conv2d(c, sd[f"fc{i}.weight"], sd[f"fc{i}.bias"]))
    c1 = outs[0]
    toks = [o.flatten(2).transpose(1, 2) for o in outs[1:]]
    return c1, toks[0], toks[1], toks[2]


def feature_decoder(sd, x, relu_masks=None):
    for k in (1, 2, 3, 4):
        x = _cbr(sd, f"decoder_{k}.0.", f"decoder_{k}.1.", x, 1, 1, relu_masks)
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(x, sd["final_out.weight"], sd["final_out.bias"], padding=1)


def dice_loss(prob, target):
    """prob: output of Softmax(1) on the logits [B,K,H,W]; target [B,H,W] int64."""
    p = torch.softmax(prob, 1)                                   # second softmax (segloss/dice.py:23)
    onehot = torch.zeros_like(p).scatter_(1, target.unsqueeze(1).long(), 1.0)
    inter = (p * onehot).sum((2, 3))
    dice = 2 * inter / (p.sum((2, 3)) + onehot.sum((2, 3)) + 10e-20)
    return 1.0 - dice.mean()


def adapter_encoder(vit_sd, spm_sd, inj_sd, ext_sd, img, num_heads, patch=14, n_points=4):
    """Returns dict with the three 42x42 maps the decoder concatenates plus the adapter state."""
    B, _, H, W = img.shape
    C = vit_sd["cls_token"].shape[-1]
    depth = vit.depth_of(vit_sd)
    d1, d2 = adapter.deform_inputs(H, W, patch)
    # reference points are fp32 cell centres (adapter_blocks.py:9-22); same values on the input's device / dtype
    d1 = [d1[0].to(img.device, img.dtype), d1[1], d1[2]]
    d2 = [d2[0].to(img.device, img.dtype), d2[1], d2[2]]
    Hc, Wc = H // 16, W // 16

    c1, c2, c3, c4 = spm(spm_sd, img)
    c = torch.cat([c2, c3, c4], 1)                               # level_embed is zeros (train.py:277)

    with torch.no_grad():
        taps = vit.get_intermediate_layers(vit_sd, img, 4, num_heads, patch, return_class_token=True)
    taps = [t for t, _ in taps]                                  # [last_4, last_3, last_2, last]

    x = vit.patch_embed(vit_sd, "patch_embed.", img, patch)      # no CLS, no pos-embed (train.py:300)
    for i in range(depth - 3):
        x = layers.block(vit_sd, f"blocks.{i}.", x, num_heads)

    for stage in range(4):
        if stage > 0:
            x = layers.block(vit_sd, f"blocks.{depth - 4 + stage}.", x, num_heads)
        x = adapter.cavit(inj_sd, "", x, d1[0], c, d1[1], n_levels=3, n_points=n_points)
        c = adapter.cacnn(ext_sd, "", c, d2[0], x, d2[1], Hc, Wc, n_levels=1, n_points=n_points)
        x = x + taps[stage]

    gh, gw = H // patch, W // patch
    out_last = x.transpose(1, 2).reshape(B, C, gh, gw)
    out_vit = taps[3].transpose(1, 2).reshape(B, C, gh, gw)
    s4 = int(round(c4.shape[1] ** 0.5))
    c4m = c4.transpose(1, 2).reshape(B, C, s4, s4)
    dy, dx = gh - s4, gw - s4
    c4m = F.pad(c4m, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return {"x": x, "c": c, "feat": torch.cat([out_last, c4m, out_vit], 1)}
