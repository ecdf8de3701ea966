"""MultiScaleAttention / MultiscaleBlock (MyModel path).

Mirrors models/multiscale_attn.py:135-304 of the reference: a timm `Attention` with a second
QKV/proj set (`qkv1`/`proj1`, used for scale attention) beside the inherited `qkv`/`proj` (used
for region attention), softmax scale 2*dim**-0.5 (:142), inside a timm `Block` layout with
LayerScale.  Parameter containers only; the math runs in the sm_100a kernels.
"""
from __future__ import annotations

from functools import partial
from typing import Dict

import torch
from torch import nn

from . import engine
from .vit_layout import AttentionParams, LayerScale, Mlp


class MultiScaleAttention(AttentionParams):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0, proj_drop=0):
        # positional call of the reference (multiscale_attn.py:137): attn_drop -> qk_norm (App. A D10)
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        self.scale = 2 * dim**-0.5
        self.qkv1 = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop1 = nn.Dropout(attn_drop)
        self.proj1 = nn.Linear(dim, dim)
        self.proj_drop1 = nn.Dropout(proj_drop)


class MultiscaleBlock(nn.Module, engine.PackCache):
    """timm Block layout (norm1, attn, ls1, drop_path1, norm2, mlp, ls2, drop_path2) with
    `attn` replaced by MultiScaleAttention (multiscale_attn.py:224-262)."""

    def __init__(self, dim, num_heads, mlp_ratio=4, qkv_bias=False, qk_norm=False, init_values=None,
                 proj_drop=0, attn_drop=0, drop_path=0, norm_layer=nn.LayerNorm, act_layer=nn.GELU):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = MultiScaleAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop,
                                        proj_drop=proj_drop)
        self.ls1 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), drop=proj_drop)
        self.ls2 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = nn.Identity()

    def pack(self, precision: str) -> Dict:
        def build():
            pl = partial(engine.pack_linear, precision=precision)
            at, mlp = self.attn, self.mlp
            # scale attention uses the second weight set (forward_with_scale :149-166)
            d = engine.pack_scale_block(
                precision, self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias,
                (at.qkv1.weight, at.qkv1.bias), (at.proj1.weight, at.proj1.bias),
                (mlp.fc1.weight, mlp.fc1.bias), (mlp.fc2.weight, mlp.fc2.bias),
                self.ls1.gamma if isinstance(self.ls1, LayerScale) else None,
                self.ls2.gamma if isinstance(self.ls2, LayerScale) else None)
            # region attention uses the inherited set (forward_with_region :190-221)
            d["region"] = {"qkv": pl(at.qkv.weight, at.qkv.bias), "proj": pl(at.proj.weight, at.proj.bias)}
            return d

        return self.packed(build, self, precision)
