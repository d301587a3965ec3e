"""Oracle: DINOv2 transformer building blocks (test infrastructure only).

Functional restatements over a reference ``state_dict``:

  * ``layer_norm``   <- ``nn.LayerNorm(eps=1e-6)``   (dinov2/models/vision_transformer.py:89)
  * ``attention``    <- ``Attention.forward`` naive path (dinov2/layers/attention.py:56-69);
                        ``MemEffAttention`` falls back to it without xformers (:73-77)
  * ``mlp``          <- ``Mlp.forward`` exact-erf GELU (dinov2/layers/mlp.py:34-40)
  * ``layer_scale``  <- ``LayerScale.forward``       (dinov2/layers/layer_scale.py:26-27)
  * ``block``        <- ``Block.forward`` eval path   (dinov2/layers/block.py:89-114)
"""
import torch
import torch.nn.functional as F

LN_EPS = 1e-6


def layer_norm(sd, prefix, x, eps=LN_EPS):
    w = sd[prefix + "weight"]
    b = sd[prefix + "bias"]
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def attention(sd, prefix, x, num_heads):
    B, T, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd[prefix + "qkv.weight"], sd.get(prefix + "qkv.bias"))
    qkv = qkv.view(B, T, 3, num_heads, hd)
    q = qkv[:, :, 0].transpose(1, 2) * hd ** -0.5       # q is pre-scaled (:60)
    k = qkv[:, :, 1].transpose(1, 2)
    v = qkv[:, :, 2].transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, T, C)
    return F.linear(o, sd[prefix + "proj.weight"], sd.get(prefix + "proj.bias"))


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x * 0.7071067811865476))


def mlp(sd, prefix, x):
    h = F.linear(x, sd[prefix + "fc1.weight"], sd.get(prefix + "fc1.bias"))
    h = gelu_erf(h)
    return F.linear(h, sd[prefix + "fc2.weight"], sd.get(prefix + "fc2.bias"))


def layer_scale(sd, prefix, x):
    key = prefix + "gamma"
    return x * sd[key] if key in sd else x          # init_values=None -> nn.Identity (block.py:72)


def block(sd, prefix, x, num_heads):
    x = x + layer_scale(sd, prefix + "ls1.", attention(sd, prefix + "attn.", layer_norm(sd, prefix + "norm1.", x), num_heads))
    x = x + layer_scale(sd, prefix + "ls2.", mlp(sd, prefix + "mlp.", layer_norm(sd, prefix + "norm2.", x)))
    return x
