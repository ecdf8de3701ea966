"""The "library bar" of BASELINE.md §5: the same DuoFormer forward as plain torch ops (cuBLAS / cuDNN / SDPA, eager)
on the same GPU and the same parameters — what a user gets from the reference's PyTorch code on a B200 without this
package.  Measurement infrastructure only (bench.py's `library_bar` key, tests/test_library_bar_gpu.py); never
imported by the package, and it calls none of the package's kernels.

Follows the reference forward (model_wo_extra_params.py:226-302, scale_attention.py:28-45, 90-93, 183-211, 330-344)
for the learned-scale-token ("random") configuration of the bench.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F


class EagerDuoFormer:
    def __init__(self, model: torch.nn.Module, dtype: torch.dtype = torch.bfloat16):
        from duoformer_tcga_b200.index_tables import num_scale_tokens, stages_used, token_row_maps

        assert getattr(model, "scale_token", "random") != "channel", "library bar covers the learned scale token"
        self.dtype = dtype
        m = copy.deepcopy(model).eval().to(dtype)
        self.trunk = m.resnet_projector.to(memory_format=torch.channels_last)
        self.by_scale = model.backbone == "r50_Swav"
        self.proj = m.projection
        self.vt = m.vision_transformer
        self.channel_token = m.channel_token
        self.num_layers = model.num_layers
        self.S = num_scale_tokens(self.num_layers)
        self.stages = stages_used(self.num_layers)
        self._row_maps = token_row_maps
        self._maps = {}

    def _features(self, x):
        if self.by_scale:
            return {i: o for i, o in enumerate(self.trunk(x))}
        feats = {}
        for name, mod in self.trunk.named_children():
            x = mod(x)
            if name in ("4", "5", "6", "7"):
                feats[int(name) - 4] = x
        return feats

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        vt, H = self.vt, self.vt.num_heads
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        feats = self._features(x)
        B, g = x.shape[0], feats[3].shape[-1]
        P, S, D = g * g, self.S, vt.embed_dim
        if g not in self._maps:
            self._maps[g] = {k: v.to(x.device).long() for k, v in self._row_maps(self.num_layers, g).items()}
        X = torch.empty(B, P * S, D, dtype=self.dtype, device=x.device)
        X[:, 0::S] = self.channel_token.reshape(1, 1, D)
        for k in self.stages:
            y = self.proj.head(k)(feats[k])  # 1x1 conv
            X[:, self._maps[g][k]] = y.flatten(2).transpose(1, 2)
        X = X.view(B, P, S, D) + vt.pos_embed_for_scale
        for blk in vt.scaleBlocks:
            h = F.layer_norm(X, (D,), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
            qkv = F.linear(h, blk.attn.qkv.weight, blk.attn.qkv.bias).view(B, P, S, 3, H, D // H).permute(3, 0, 1, 4, 2, 5)
            a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], scale=blk.attn.scale)
            X = X + F.linear(a.transpose(2, 3).reshape(B, P, S, D), blk.attn.proj.weight, blk.attn.proj.bias)
            h = F.layer_norm(X, (D,), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
            h = F.gelu(F.linear(h, blk.mlp.fc1.weight, blk.mlp.fc1.bias))
            X = X + F.linear(h, blk.mlp.fc2.weight, blk.mlp.fc2.bias)
        z = torch.cat((vt.cls_token.expand(B, -1, -1), X[:, :, 0, :]), dim=1) + vt.pos_embed
        N = P + 1
        for blk in vt.blocks:
            qkv = F.linear(z, blk.attn.qkv.weight, blk.attn.qkv.bias).view(B, N, 3, H, D // H).permute(2, 0, 3, 1, 4)
            a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], scale=blk.attn.scale)
            z = F.linear(a.transpose(1, 2).reshape(B, N, D), blk.attn.proj.weight, blk.attn.proj.bias)
        return F.linear(z[:, 0], vt.head.weight, vt.head.bias).float()


def measure(model: torch.nn.Module, batch: int = 64, iters: int = 3, size: int = 224, dtype: torch.dtype = torch.bfloat16):
    """images/s of the eager library forward at `batch` (CUDA events, 2 warm-up passes)."""
    dev = next(model.parameters()).device
    eager = EagerDuoFormer(model, dtype)
    x = torch.randn(batch, 3, size, size, device=dev)
    for _ in range(2):
        y = eager(x)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = eager(x)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    return {"value": batch / ms * 1000.0, "unit": "images/s", "batch": batch, "ms_per_step": ms,
            "what": "the same forward as plain torch ops on this GPU (cuDNN trunk, cuBLAS Linears, SDPA attention, "
                    f"eager, {str(dtype).replace('torch.', '')}), same parameters",
            "finite": bool(torch.isfinite(y).all())}, eager
