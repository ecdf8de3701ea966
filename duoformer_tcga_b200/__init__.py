"""duoformer_tcga_b200 — B200-native (sm_100a) DuoFormer multi-scale transformer forward path.

Drop-in for the reference's `models` package on this path: `build_model`,
`build_model_no_extra_params`, `MyModel`, `MyModel_no_extra_params` keep the reference's
signatures (models/__init__.py:12-70) and state_dict schema; compute runs in
libduoformer_sm100.so (hand-written tcgen05/TMA/TMEM + CUDA-core kernels) through a C ABI.
"""
from .model import MyModel, count_parameters  # noqa: F401
from .model_wo_extra_params import MyModel_no_extra_params  # noqa: F401
from .multi_vision_transformer import MultiscaleTransformer  # noqa: F401
from .multiscale_attn import MultiScaleAttention, MultiscaleBlock  # noqa: F401
from .projection_head import (Channel_Projector_All, Channel_Projector_layer1, Channel_Projector_layer2,  # noqa: F401
                              Channel_Projector_layer3, Projection)
from .resnet50ssl import ResNetTrunk, ResNetTrunkByScale, resnet50FeatureExtractor  # noqa: F401
from .scale_attention import (AttentionForPatch, AttentionForScale, MultiscaleFormer, PatchBlock,  # noqa: F401
                              ScaleBlock)

__version__ = "0.1.0"


def build_model(depth=12, patch_size=49, embed_dim=256, num_heads=6, init_values=1e-5, num_classes=100,
                num_layers=4, proj_dim=384, model_ver="scaleformer", pretrained=True, freeze=True):
    """Factory of the "with extra params" DuoFormer; signature of the reference's models/__init__.py:12-24
    (positional order and defaults kept), every argument forwarded by keyword (:25-37)."""
    return MyModel(**dict(locals()))


def build_model_no_extra_params(depth=12, embed_dim=256, num_heads=6, num_classes=100, num_layers=4,
                                num_patches=49, proj_dim=384, mlp_ratio=4.0, attn_drop_rate=0.0,
                                proj_drop_rate=0.0, freeze_backbone=True, backbone="r50", pretrained=True):
    """Factory of the from-scratch MultiscaleFormer DuoFormer; signature of models/__init__.py:40-54.
    The reference forwards `pretrained=` to a class that does not accept it (:53,:69, App. A D2);
    here it is honoured (False = random-init trunk, no network)."""
    return MyModel_no_extra_params(**dict(locals()))
