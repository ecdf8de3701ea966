"""Channel-token branch on the tcgen05 GEMM.

Reference: `Channel_Projector_layer1/2/3` + `Channel_Projector_All` (projection_head.py:152-268),
called from model.py:279-289 and model_wo_extra_params.py:236-248:

    c0 = maxpool2(conv3x3s2(conv3x3s2(F0)))   256 ch, 8g -> 4g -> 2g -> g
    c1 = maxpool2(conv3x3s2(F1))               512 ch, 4g -> 2g -> g
    c2 = maxpool2(F2)                          1024 ch
    cat(c0, c1, c2, F3) -> [B, 3840, g, g] -> 4 x ReLU(BN(conv3x3(.))) -> [B, 768, g*g]

Every 3x3 convolution is an IMPLICIT GEMM on the tcgen05 convolution kernel (`duo_conv2d`, csrc/conv_tcgen05.cu: the
K loop walks the nine taps, each operand tile is one 4-D TMA box load of the NHWC map — the [B*49, 34 560] im2col matrix
of the 3840 -> 768 convolution is never written) with the weight permuted to [N, ky, kx, c], BatchNorm folded in and
ReLU in the epilogue.  The stage maps are consumed in the trunk's 16-bit type (fp16 operands), every output is bf16
(un-normalised sums: bf16's range).  Only the LAST convolution, whose result is the fp32 scale token, runs as
`duo_im2col3x3` + `duo_gemm` with the fp32 epilogue (768 input channels: a [B*49, 6 912] matrix).  The max-pools and the
concatenation are `duo_pool_to_slice` writes into channel slices of one NHWC buffer.  Used in the bf16 mode; the
fp32-accuracy mode keeps the fp32 cuDNN modules.
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn

from . import engine, ops


def _conv_as_gemm_weight(w: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """[N, C, 3, 3] -> 16-bit [N, 9*C] with column order (ky, kx, c) — the tap order of duo_conv2d / duo_im2col3x3."""
    return w.detach().float().permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(dtype).contiguous()


def _both(w: torch.Tensor) -> Dict[torch.dtype, torch.Tensor]:
    """Weights of a convolution fed by a trunk map: one copy per 16-bit type the trunk may deliver."""
    return {dt: _conv_as_gemm_weight(w, dt) for dt in (torch.float16, torch.bfloat16)}


def _fold_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = conv.weight.detach().float() * s[:, None, None, None]
    b = (conv.bias.detach().float() - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return w, b


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    return t.permute(0, 2, 3, 1).contiguous()  # free when t is channels-last


class ChannelBranch(engine.PackCache):
    def __init__(self, proj1: nn.Module, proj2: nn.Module, proj_all: nn.Module):
        self.proj1, self.proj2, self.proj_all = proj1, proj2, proj_all
        self._holder = nn.ModuleList([proj1, proj2, proj_all])  # one parameter signature for the cache

    def pack(self) -> Dict:
        def build():
            p = {
                "c11": (_both(self.proj1.conv1.weight), engine._f32(self.proj1.conv1.bias)),
                "c12": (_both(self.proj1.conv2.weight), engine._f32(self.proj1.conv2.bias)),
                "c21": (_both(self.proj2.conv1.weight), engine._f32(self.proj2.conv1.bias)),
                "all": [],
            }
            for cb in self.proj_all.nConvs:
                w, b = _fold_bn(cb.conv, cb.norm)
                p["all"].append((_conv_as_gemm_weight(w), b.contiguous()))
            return p

        return self.packed(build, self._holder, "bf16")

    @torch.no_grad()
    def __call__(self, feats: Dict[int, torch.Tensor]) -> torch.Tensor:
        """feats[k]: [B, C_k, g*2^(3-k), g*2^(3-k)] (any float dtype).  Returns fp32 [B, g*g, 768]."""
        pk = self.pack()
        f0, f1, f2, f3 = (_nhwc(feats[k]) for k in range(4))
        B, g = f3.shape[0], f3.shape[1]
        dev = f3.device
        bf = dict(dtype=torch.bfloat16, device=dev)

        def conv(x_nhwc, wb, stride, relu=False, out_f32=False):
            w = wb[0][x_nhwc.dtype] if isinstance(wb[0], dict) else wb[0]
            if not out_f32 and x_nhwc.dtype in (torch.float16, torch.bfloat16):
                return ops.conv2d(x_nhwc, w, wb[1], 3, stride, relu, out_dtype=torch.bfloat16)
            Bx, H, W, _ = x_nhwc.shape
            Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
            a = ops.im2col3x3(x_nhwc, stride)
            if isinstance(wb[0], dict):
                w = wb[0][torch.bfloat16]
            n = w.shape[0]
            out = torch.empty(Bx * Ho * Wo, n, dtype=torch.float32 if out_f32 else torch.bfloat16, device=dev)
            ops.gemm(a, w, wb[1], out, ops.EPI_F32 if out_f32 else ops.EPI_BF16, relu=relu)
            return out.view(Bx, Ho, Wo, n)

        c0, c1, c2, c3 = f0.shape[3], f1.shape[3], f2.shape[3], f3.shape[3]
        cat = torch.empty(B * g * g, c0 + c1 + c2 + c3, **bf)
        y = conv(conv(f0, pk["c11"], 2), pk["c12"], 2)           # [B, 2g, 2g, 256]
        ops.pool_to_slice(y, cat[:, 0:c0], 2)
        y = conv(f1, pk["c21"], 2)                                # [B, 2g, 2g, 512]
        ops.pool_to_slice(y, cat[:, c0:c0 + c1], 2)
        ops.pool_to_slice(f2, cat[:, c0 + c1:c0 + c1 + c2], 2)
        ops.pool_to_slice(f3, cat[:, c0 + c1 + c2:], 1)
        x = cat.view(B, g, g, -1)
        n_all = len(pk["all"])
        for i, wb in enumerate(pk["all"]):
            x = conv(x, wb, 1, relu=True, out_f32=(i == n_all - 1))
        return x.view(B, g * g, -1)
