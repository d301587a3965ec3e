"""CPU-only host-logic tests: constructor/API/state_dict compatibility with the reference modules
(keys taken from the golden fixtures the reference produced) and reference error behaviour."""
import warnings

import pytest
import torch

import adaptersis_b200 as asis
from oracle import msda as o_msda


def test_state_dict_keys_match_reference(golden):
    g = golden("block.pt")
    blk = asis.Block(dim=64, num_heads=4, mlp_ratio=4.0, qkv_bias=True, proj_bias=True, ffn_bias=True, init_values=1e-5,
                     attn_class=asis.MemEffAttention)
    blk.load_state_dict(g["sd"], strict=True)
    g = golden("vit.pt")
    cfg = g["cfg"]
    vit = asis.DinoVisionTransformer(img_size=cfg["img_size"], patch_size=14, embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                                     num_heads=cfg["num_heads"], init_values=1e-5, block_chunks=0)
    vit.load_state_dict(g["sd"], strict=True)
    assert set(vit.state_dict()) == set(g["sd"])
    g = golden("adapter.pt")
    asis.CAViT(dim=32, n_levels=3, num_heads=4, n_points=4).load_state_dict(g["inj_sd"], strict=True)
    asis.CACNN(dim=32, n_levels=1, num_heads=4, n_points=4, cffn_ratio=0.25).load_state_dict(g["ext_sd"], strict=True)
    g = golden("msda_module.pt")
    asis.MSDeformAttn(**g["cfg"]).load_state_dict(g["sd"], strict=True)
    g = golden("encoder.pt")
    cfg = g["cfg"]
    asis.FeatureEncoder(inplanes=cfg["inplanes"], embed_dim=cfg["dim"]).load_state_dict(g["spm_sd"], strict=True)
    asis.FeatureDecoder(embed_dim=cfg["dim"], num_classes=2, features=cfg["dec_features"]).load_state_dict(g["dec_sd"], strict=True)


def test_mask_transformer_keys_match_reference(golden):
    # config[3]: same constructor call as eval/eval_dinov2_masktrans.py:136-139 (scaled down), same state_dict keys / shapes
    from adaptersis_b200.masktrans import MaskTransformer
    g = golden("masktrans.pt")
    c = g["cfg"]
    m = MaskTransformer(n_cls=c["n_cls"], patch_size=14, d_encoder=c["d"], n_layers=2, n_heads=c["heads"], d_model=c["d"],
                        d_ff=4 * c["d"], drop_path_rate=0.0, dropout=0.1)
    res = m.load_state_dict(g["sd"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert m.no_weight_decay() == {"cls_emb"}
    import pytest
    with pytest.raises(NotImplementedError):
        m.get_attention_map(g["x"], 0)


def test_vit_factories_match_reference_sizes():
    # parameter counts probed from the reference: block 12,598,272; ViT-L 304,368,640
    blk = asis.Block(dim=1024, num_heads=16, qkv_bias=True, init_values=1e-5, attn_class=asis.MemEffAttention)
    assert sum(p.numel() for p in blk.parameters()) == 12_598_272
    with torch.device("meta"):
        m = asis.vit_large(patch_size=14, img_size=518, init_values=1e-5, block_chunks=0)
    assert sum(p.numel() for p in m.parameters()) == 304_368_640
    assert sum(p.numel() for p in asis.CAViT(dim=1024, n_levels=3, num_heads=8, n_points=4).parameters()) == 2_399_520
    assert sum(p.numel() for p in asis.CACNN(dim=1024, n_levels=1, num_heads=8, n_points=4, cffn_ratio=0.25).parameters()) == 2_731_872


def test_msda_constructor_contract(golden):
    with pytest.raises(ValueError, match="divisible"):
        asis.MSDeformAttn(d_model=30, n_heads=4)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        asis.MSDeformAttn(d_model=96, n_levels=1, n_heads=8, n_points=2)     # head dim 12: warning, not error
        assert any("power of 2" in str(x.message) for x in w)
    g = golden("msda_module.pt")
    m = asis.MSDeformAttn(**g["cfg"])
    assert torch.equal(m.sampling_offsets.bias.detach(), g["default_bias"])
    assert torch.equal(m.sampling_offsets.bias.detach(), o_msda.msda_reset_bias(4, 3, 4))
    assert float(m.attention_weights.weight.abs().sum()) == 0.0 and m.im2col_step == 64


def test_deform_inputs_match_reference(golden):
    g = golden("adapter.pt")
    d1, d2 = asis.deform_inputs(torch.zeros(1, 3, 588, 588), 14)
    for a, b in zip(d1 + d2, g["d1"] + g["d2"]):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    assert d1[1].tolist() == [[73, 73], [36, 36], [18, 18]] and d1[2].tolist() == [0, 5329, 6625]


def test_default_and_chunked_block_layouts_match_reference_keys():
    """the reference's constructor default is block_chunks=1 (FSDP chunk layout, keys
    ``blocks.<chunk>.<index>.*``); it must construct, and both layouts must carry the reference's keys."""
    m = asis.DinoVisionTransformer(img_size=28, patch_size=14, embed_dim=32, depth=4, num_heads=2)   # all defaults
    assert m.chunked_blocks and "blocks.0.3.norm1.weight" in m.state_dict()
    m2 = asis.DinoVisionTransformer(img_size=28, patch_size=14, embed_dim=32, depth=4, num_heads=2, block_chunks=2)
    assert "blocks.1.2.norm1.weight" in m2.state_dict() and "blocks.0.2.norm1.weight" not in m2.state_dict()
    assert len(m2._real_blocks()) == 4
    flat = asis.DinoVisionTransformer(img_size=28, patch_size=14, embed_dim=32, depth=4, num_heads=2, block_chunks=0)
    assert "blocks.3.norm1.weight" in flat.state_dict()
    with pytest.raises(ValueError, match="block_chunks=0"):
        asis.AdapterEncoder(model=m2)


def test_unsupported_configs_are_rejected():
    with pytest.raises(NotImplementedError):
        asis.DinoVisionTransformer(patch_size=14, block_chunks=0, drop_path_rate=0.1)
    blk = asis.NestedTensorBlock(dim=32, num_heads=2)
    with pytest.raises(AssertionError, match="xFormers"):
        blk([torch.zeros(1, 2, 32)])
    att = asis.MemEffAttention(32, 2)
    with pytest.raises(AssertionError, match="xFormers"):
        att(torch.zeros(1, 2, 32), attn_bias=object())
