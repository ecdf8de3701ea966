"""Token projection (1x1 convs) and the channel-token branch.

Mirrors models/projection_head.py of the reference: `Projection` :11-149 (per-stage 1x1 Conv2d
C_k -> proj_dim, kaiming-normal weight / N(0,1e-6) bias), `Channel_Projector_layer1/2/3`
:152-222, `ConvBatchNorm` :242-254, `Channel_Projector_All` :257-268.

On the hot path `Projection` is never run as a convolution: its weights feed the fused
tcgen05 GEMM + token-scatter kernel (see token_builder.py).  `Projection.forward` is kept for
API parity and runs the same GEMM kernel with a plain fp32 epilogue.  The channel-token branch
(3x3 convs + BN + ReLU + max-pools) is declared here for the state_dict schema; in bf16 mode it runs as
implicit-GEMM convolutions on tcgen05 + pool-to-slice kernels (channel_branch.py), in fp32 mode through
these torch modules (fp32 cuDNN, TF32 off).
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn

from . import engine, ops
from .index_tables import STAGE_CHANNELS_R18, STAGE_CHANNELS_R50
from .vit_layout import kaiming_conv_init_


class Projection(nn.Module):
    def __init__(self, num_layers=2, proj_dim=768, backbone="r50"):
        super().__init__()
        self.num_layers = num_layers
        self.proj_dim = proj_dim
        self.backbone = backbone
        ch = STAGE_CHANNELS_R50 if backbone == "r50" else STAGE_CHANNELS_R18
        if backbone not in ("r50", "r18"):
            raise ValueError(f"unknown backbone {backbone!r}")
        if num_layers == 1:
            self.proj_heads = nn.Conv2d(ch[3], proj_dim, kernel_size=(1, 1), stride=(1, 1))
            kaiming_conv_init_(self.proj_heads)
        else:
            if backbone == "r18" and num_layers in (2, 3):
                # projection_head.py:66-93 builds heads {2,1} / {0,2,1} that the forward never matches
                raise NotImplementedError("r18 with 2 or 3 scales is broken in the reference (App. A D9)")
            for k in [3, 2, 1, 0][:num_layers]:
                conv = nn.Conv2d(ch[k], proj_dim, kernel_size=(1, 1), stride=(1, 1))
                kaiming_conv_init_(conv)
                setattr(self, f"proj_heads{k}", conv)

    def head(self, k: int) -> nn.Conv2d:
        return self.proj_heads if self.num_layers == 1 else getattr(self, f"proj_heads{k}")

    @torch.no_grad()
    def forward(self, x: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """{stage: [B,C_k,h,w]} -> {stage: [B,proj_dim,h,w]} (projection_head.py:134-149)."""
        out = {}
        for key, fea in x.items():
            engine.require_cuda(fea, "Projection.forward")
            conv = self.head(int(key))
            B, C, H, W = fea.shape
            a = fea.to(torch.float32).permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
            A = torch.empty(B * H * W, C, dtype=torch.bfloat16, device=fea.device)
            ops.convert(a, A)
            w, b = engine.pack_linear(conv.weight, conv.bias, "bf16")
            y = torch.empty(B * H * W, self.proj_dim, dtype=torch.float32, device=fea.device)
            ops.gemm(A, w, b, y, ops.EPI_F32)
            out[key] = y.view(B, H, W, self.proj_dim).permute(0, 3, 1, 2)
        return out


class Channel_Projector_layer1(nn.Module):
    """Stage 0: two stride-2 3x3 convs (no activation) + 2x2 max-pool, 56 -> 7 (:152-184)."""

    def __init__(self, backbone="r50"):
        super().__init__()
        c = 256 if backbone == "r50" else 64
        self.conv1 = nn.Conv2d(c, c, kernel_size=3, stride=2, padding=1)
        self.conv2 = nn.Conv2d(c, c, kernel_size=3, stride=2, padding=1)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        kaiming_conv_init_(self.conv1)
        kaiming_conv_init_(self.conv2)

    def forward(self, x):
        return self.pool(self.conv2(self.conv1(x)))


class Channel_Projector_layer2(nn.Module):
    """Stage 1: one stride-2 3x3 conv + 2x2 max-pool, 28 -> 7 (:187-212)."""

    def __init__(self, backbone="r50"):
        super().__init__()
        c = 512 if backbone == "r50" else 128
        self.conv1 = nn.Conv2d(c, c, kernel_size=3, stride=2, padding=1)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        kaiming_conv_init_(self.conv1)

    def forward(self, x):
        return self.pool(self.conv1(x))


class Channel_Projector_layer3(nn.Module):
    """Stage 2: 2x2 max-pool, 14 -> 7 (:215-222)."""

    def __init__(self):
        super().__init__()
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)

    def forward(self, x):
        return self.pool(x)


class ConvBatchNorm(nn.Module):
    """(convolution => [BN] => ReLU) (:242-254)."""

    def __init__(self, in_channels, out_channels, activation="ReLU"):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.norm = nn.BatchNorm2d(out_channels)
        self.activation = nn.ReLU()

    def forward(self, x):
        return self.activation(self.norm(self.conv(x)))


def _make_nConv(in_channels, out_channels, nb_Conv, activation="ReLU"):
    layers = [ConvBatchNorm(in_channels, out_channels, activation)]
    for _ in range(nb_Conv - 1):
        layers.append(ConvBatchNorm(out_channels, out_channels, activation))
    return nn.Sequential(*layers)


class Channel_Projector_All(nn.Module):
    """4 x (conv3x3 + BN + ReLU), 3840 -> 768, flattened to [B, 768, P] (:257-268)."""

    def __init__(self, backbone="r50"):
        super().__init__()
        self.nConvs = _make_nConv(3840 if backbone == "r50" else 384, 768, 4)

    def forward(self, x):
        return torch.flatten(self.nConvs(x), start_dim=2)
