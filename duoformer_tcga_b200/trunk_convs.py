"""ResNet-50 trunk on the package's own convolution kernel (csrc/conv_tcgen05.cu) — no cuDNN.

Producer of the stage maps the token builder consumes: `get_features` model_wo_extra_params.py:214-224 /
model.py:213-223 and `ResNetTrunkByScale.forward` resnet50ssl.py:35-45 (torchvision ResNet: conv1 7x7/2 + bn1 + ReLU,
MaxPool2d(3, 2, 1), four stages of Bottlenecks).  Built from the BN-folded fp32 copy of the trunk (token_builder.
_fold_batchnorm_): every convolution becomes one implicit-GEMM launch with bias / residual add / ReLU in its epilogue,
activations stay NHWC 16-bit (fp16 by default) from the packed image to the four taps:

  stem        ops.stem_pack (image -> padded row-pair tensor, the fp16 range-guard factor folded in) + ops.stem_conv7x7 + ReLU
  max-pool    ops.maxpool3x3s2
  Bottleneck  conv1 1x1 + ReLU;  conv2 3x3 stride s + ReLU;  conv3 1x1 + bias + identity + ReLU (the identity tile is
              TMA-loaded into the epilogue's staging slot), or — first block of a stage — conv3 with the downsample
              1x1 / stride s convolution of the block input FUSED: its K blocks are appended to conv3's K loop (second
              input tensor map, concatenated weights, summed biases), so the shortcut tensor is never written.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops


def _pair(v) -> Tuple[int, int]:
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _plain_conv(conv, ksize: int, strides=(1, 2)) -> bool:
    return (isinstance(conv, nn.Conv2d) and _pair(conv.kernel_size) == (ksize, ksize) and conv.groups == 1
            and _pair(conv.dilation) == (1, 1) and _pair(conv.padding) == (ksize // 2, ksize // 2)
            and _pair(conv.stride)[0] == _pair(conv.stride)[1] and _pair(conv.stride)[0] in strides
            and conv.in_channels % 64 == 0 and conv.out_channels % 64 == 0 and conv.padding_mode == "zeros")


def _gemm_weight(conv: nn.Conv2d) -> torch.Tensor:
    """[Cout, Cin, k, k] -> fp32 [Cout, k*k*Cin] in (ky, kx, c) column order."""
    w = conv.weight.detach().float()
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


class _Conv:
    """One packed convolution: weight [Cout, k*k*Cin (+ Cin2)] 16-bit, bias fp32 or None; `shortcut` = a 1x1 convolution
    of a second tensor fused into the K loop (the Bottleneck's downsample)."""

    def __init__(self, conv: nn.Conv2d, dtype: torch.dtype, bias: Optional[torch.Tensor], shortcut: Optional[nn.Conv2d] = None):
        self.ksize = int(conv.kernel_size[0])
        self.stride = int(_pair(conv.stride)[0])
        w = _gemm_weight(conv)
        self.stride2 = 1
        if shortcut is not None:
            w = torch.cat([w, _gemm_weight(shortcut)], dim=1)
            self.stride2 = int(_pair(shortcut.stride)[0])
        self.weight = w.to(dtype).contiguous()
        self.bias = None if bias is None else bias.detach().float().contiguous()

    def __call__(self, x: torch.Tensor, relu: bool, residual: Optional[torch.Tensor] = None,
                 in2: Optional[torch.Tensor] = None) -> torch.Tensor:
        return ops.conv2d(x, self.weight, self.bias, self.ksize, self.stride, relu, residual, in2=in2, stride2=self.stride2)


class _Bottleneck:
    def __init__(self, blk: nn.Module, dtype: torch.dtype):
        self.c1 = _Conv(blk.conv1, dtype, blk.conv1.bias)
        self.c2 = _Conv(blk.conv2, dtype, blk.conv2.bias)
        self.fused_shortcut = blk.downsample is not None
        if self.fused_shortcut:  # relu(conv3(.) + b3 + conv_ds(x) + b_ds): one accumulator, one bias
            ds = blk.downsample[0]
            self.c3 = _Conv(blk.conv3, dtype, blk.conv3.bias.detach().float() + ds.bias.detach().float(), shortcut=ds)
        else:
            self.c3 = _Conv(blk.conv3, dtype, blk.conv3.bias)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        o = self.c1(x, True)
        o = self.c2(o, True)
        return self.c3(o, True, in2=x) if self.fused_shortcut else self.c3(o, True, residual=x)


def _parts(trunk: nn.Module, by_scale: bool):
    if by_scale:
        return trunk.conv1, trunk.maxpool, [trunk.layer1, trunk.layer2, trunk.layer3, trunk.layer4]
    ch = dict(trunk.named_children())
    return ch["0"], ch["3"], [ch["4"], ch["5"], ch["6"], ch["7"]]


def eligible(folded: nn.Module, by_scale: bool) -> bool:
    """True for a BN-folded torchvision ResNet-50-style trunk: 7x7/2/3 stem on 3 channels, MaxPool2d(3, 2, 1), stages of
    Bottlenecks whose convolutions are plain (groups 1, dilation 1, 64-multiple channels, stride 1 or 2)."""
    try:
        stem, pool, layers = _parts(folded, by_scale)
    except (KeyError, AttributeError):
        return False
    if not (isinstance(stem, nn.Conv2d) and stem.in_channels == 3 and _pair(stem.kernel_size) == (7, 7)
            and _pair(stem.stride) == (2, 2) and _pair(stem.padding) == (3, 3) and stem.groups == 1
            and _pair(stem.dilation) == (1, 1) and stem.out_channels % 64 == 0 and stem.bias is not None):
        return False
    if not (isinstance(pool, nn.MaxPool2d) and _pair(pool.kernel_size) == (3, 3) and _pair(pool.stride) == (2, 2)
            and _pair(pool.padding) == (1, 1) and _pair(pool.dilation) == (1, 1) and not pool.ceil_mode):
        return False
    for layer in layers:
        for blk in layer:
            if blk.__class__.__name__ != "Bottleneck":
                return False
            if not (_plain_conv(blk.conv1, 1, (1,)) and _plain_conv(blk.conv2, 3) and _plain_conv(blk.conv3, 1, (1,))):
                return False
            if any(c.bias is None for c in (blk.conv1, blk.conv2, blk.conv3)):  # BN not folded
                return False
            if any(isinstance(getattr(blk, n, None), nn.BatchNorm2d) for n in ("bn1", "bn2", "bn3")):
                return False
            ds = blk.downsample
            if ds is not None:
                if not (isinstance(ds, nn.Sequential) and len(ds) == 2 and _plain_conv(ds[0], 1) and ds[0].bias is not None
                        and isinstance(ds[1], nn.Identity)):
                    return False
    return True


class OwnTrunk:
    """Packed operands of the trunk for one weight set; `folded` is the BN-folded fp32 trunk with the range-guard factor
    already applied to its biases (token_builder.TrunkRunner._packed_trunk)."""

    def __init__(self, folded: nn.Module, by_scale: bool, dtype: torch.dtype):
        assert dtype in (torch.float16, torch.bfloat16)
        stem, _, layers = _parts(folded, by_scale)
        self.dtype = dtype
        self.stem_w = ops.pack_stem_weight(stem.weight, dtype)
        self.stem_b = stem.bias.detach().float().contiguous()
        self.layers: List[List[_Bottleneck]] = [[_Bottleneck(blk, dtype) for blk in layer] for layer in layers]

    @torch.no_grad()
    def features(self, x: torch.Tensor, act_scale: float) -> Dict[int, torch.Tensor]:
        """x [B,3,H,W] (any float dtype / memory format) -> stage maps 0..3 as NCHW-shaped channels-last tensors,
        multiplied by act_scale (the biases already carry it)."""
        if x.dtype != torch.float32:
            x = x.float()
        y = ops.stem_conv7x7(ops.stem_pack(x, act_scale, self.dtype), self.stem_w, self.stem_b, relu=True)
        y = ops.maxpool3x3s2(y.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)  # NHWC view of the channels-last result
        feats: Dict[int, torch.Tensor] = {}
        for i, layer in enumerate(self.layers):
            for blk in layer:
                y = blk(y)
            feats[i] = y.permute(0, 3, 1, 2)
        return feats
