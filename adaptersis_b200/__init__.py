"""adaptersis_b200 -- B200-native (sm_100a) implementation of AdapterSIS's data-parallel hot path:
DINOv2 ViT backbone forward/backward, the injector/extractor adapter blocks and multi-scale
deformable attention, behind the reference's own module API.  All compute goes through
libasis_b200.so (include/asis_b200.h); there is no CPU fallback."""
from . import _lib  # noqa: F401
from .functional import (MSDeformAttnFunction, get_precision, invalidate_weight_cache,  # noqa: F401
                         ms_deform_attn_core, precision, set_precision)
from .ms_deform_attn import MSDeformAttn  # noqa: F401
from .adapter_blocks import CACNN, CAViT, ConvFFN, DWConv, deform_inputs, get_reference_points  # noqa: F401
from .layers import (Attention, Block, LayerScale, MemEffAttention, Mlp, NestedTensorBlock,  # noqa: F401
                     PatchEmbed)
from .vision_transformer import (DinoVisionTransformer, ModelWithIntermediateLayers,  # noqa: F401
                                 build_model_for_eval, vit_base, vit_giant2, vit_large, vit_small)
from .encoders import FeatureEncoder  # noqa: F401
from .decoders import FeatureDecoder  # noqa: F401
from .encoder import AdapterEncoder  # noqa: F401
from .masktrans import MaskTransformer  # noqa: F401
from .dp import BucketedGradAllReduce  # noqa: F401

__version__ = "0.1.0"
