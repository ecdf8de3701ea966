"""Parameter containers with the attribute layout of timm 0.9.8's ViT building blocks.

The reference sub-classes `timm.models.vision_transformer.{Attention, Block, VisionTransformer}`
and uses `timm.layers.{Mlp, DropPath}` / `LayerScale` (pinned timm==0.9.8, environmental.yml:156).
timm is not a dependency here; these classes only reproduce the *state_dict key schema* and the
initialisation those classes give so that reference checkpoints load
unchanged.  None of them computes anything: the forward math runs in the sm_100a kernels.
"""
from __future__ import annotations

import math

import torch
from torch import nn


def trunc_normal_(tensor: torch.Tensor, std: float = 1.0) -> torch.Tensor:
    # timm.layers.trunc_normal_ == torch.nn.init.trunc_normal_ with absolute bounds a=-2, b=2
    return nn.init.trunc_normal_(tensor, mean=0.0, std=std, a=-2.0, b=2.0)


def init_weights_vit(module: nn.Module) -> None:
    """timm init_weights_vit_timm: Linear -> trunc_normal(.02) weight, zero bias."""
    if isinstance(module, nn.Linear):
        trunc_normal_(module.weight, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)


def named_apply_vit_init(root: nn.Module) -> None:
    """named_apply(get_init_weights_vit(mode=""), root): depth-first over every sub-module."""
    for m in root.modules():
        init_weights_vit(m)


class Mlp(nn.Module):
    """timm.layers.Mlp layout: fc1, act, drop1, norm, fc2, drop2 (GELU is exact-erf)."""

    def __init__(self, in_features: int, hidden_features: int, drop: float = 0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.drop2 = nn.Dropout(drop)


class LayerScale(nn.Module):
    def __init__(self, dim: int, init_values: float = 1e-5):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class AttentionParams(nn.Module):
    """timm Attention(dim, num_heads, qkv_bias, qk_norm, attn_drop, proj_drop, norm_layer) layout.

    The reference calls `super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)`
    positionally (scale_attention.py:25,178; multiscale_attn.py:137), which in timm 0.9.8 shifts
    the arguments: qk_norm <- attn_drop, attn_drop <- proj_drop, proj_drop <- 0.  A non-zero
    attn_drop therefore creates (unused) q_norm / k_norm LayerNorm(head_dim) parameters — they are
    reproduced here so the state_dict matches."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, attn_drop=0.0, proj_drop=0.0,
                 norm_layer=nn.LayerNorm):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim**-0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class PatchEmbedParams(nn.Module):
    """timm PatchEmbed: only the (unused) projection conv matters for the state_dict."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


def kaiming_conv_init_(conv: nn.Conv2d) -> None:
    """projection_head.py:119-132: kaiming-normal weight, N(0, 1e-6) bias."""
    nn.init.kaiming_normal_(conv.weight)
    if conv.bias is not None:
        nn.init.normal_(conv.bias, std=1e-6)


def default_linear_init_(lin: nn.Linear) -> None:
    nn.init.kaiming_uniform_(lin.weight, a=math.sqrt(5))
    if lin.bias is not None:
        bound = 1 / math.sqrt(lin.weight.shape[1])
        nn.init.uniform_(lin.bias, -bound, bound)
