"""Deterministic synthetic weights / inputs — TEST INFRASTRUCTURE ONLY.

Weights are a pure function of (state_dict key, shape, seed): every tensor is drawn from a CPU
torch.Generator seeded by a hash of its key, so the reference model (golden generation, build
container), the oracle and the CUDA path (GPU box) all see bit-identical values without
shipping ~500 MB of weights.  Distributions follow the reference's initialisers in scale
 but give every bias / norm / running statistic a non-trivial value so that
each term of the forward is exercised, and apply the survey's "W-signal" (patch-block proj x3.3)
and "W-layerscale" (gamma ~ 0.5) adjustments so the logits depend on the input.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return torch.Generator(device="cpu").manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)


def synth_tensor(key: str, ref: torch.Tensor, seed: int) -> torch.Tensor:
    shape = tuple(ref.shape)
    g = _gen(key, seed)

    def randn(std=1.0, mean=0.0):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean

    if key.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=ref.dtype)
    leaf = key.rsplit(".", 1)[-1]
    is_trunk = key.startswith("resnet_projector.")
    # ---- BatchNorm (trunk + channel branch) ----
    if leaf == "running_mean":
        return randn(0.05)
    if leaf == "running_var":
        return 1.0 + 0.2 * torch.rand(shape, generator=g)
    is_bn = (".bn" in key or "downsample.1." in key or key.startswith("resnet_projector.1.") or
             key.startswith("resnet_projector.bn1.") or ".norm." in key and key.startswith("chann_proj_all"))
    if is_bn and leaf == "weight":
        base = 0.35 if ".bn3." in key else 1.0  # damp the residual branch: O(1) activations
        return randn(0.05, base)
    if is_bn and leaf == "bias":
        return randn(0.05)
    # ---- convolutions ----
    if len(shape) == 4:
        fan_in = shape[1] * shape[2] * shape[3]
        if key.startswith("projection."):
            return randn(math.sqrt(2.0 / fan_in))  # kaiming-normal (projection_head.py:119-132)
        if key.startswith("chann_proj"):
            return randn(math.sqrt(1.0 / fan_in))
        return randn(math.sqrt(2.0 / fan_in))
    # ---- tokens / position tables ----
    if leaf in ("channel_token", "cls_token", "pos_embed", "pos_embed_for_scale"):
        return randn(0.036 if leaf != "channel_token" else 0.3)
    if leaf == "gamma":  # LayerScale — "W-layerscale"
        return randn(0.1, 0.5)
    # ---- LayerNorm ----
    if ("norm" in key) and len(shape) == 1:
        return randn(0.1, 1.0) if leaf == "weight" else randn(0.05)
    # ---- Linear ----
    if len(shape) == 2:
        std = 0.02
        if ".blocks." in key and ".attn.proj.weight" in key:
            std = 0.02 * 3.3  # "W-signal": keep magnitude through the residual-free patch stack
        if key.endswith("head.weight"):
            std = 0.05
        return randn(std)
    if len(shape) == 1:  # biases
        return randn(0.02)
    return randn(0.02)


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Same keys/shapes/dtypes as `template`, values regenerated deterministically."""
    out = {}
    for k, v in template.items():
        t = synth_tensor(k, v, seed)
        out[k] = t.to(v.dtype) if v.is_floating_point() else t.to(v.dtype)
    return out


def synth_images(batch: int, size: int = 224, seed: int = 1234) -> torch.Tensor:
    """N(0,1) stand-in for ImageNet-normalised tiles."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(batch, 3, size, size, generator=g, dtype=torch.float32)
