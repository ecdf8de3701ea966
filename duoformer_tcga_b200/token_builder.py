"""Multiscale token builder: stage features -> tokens [B, P, S, D] (fp32, + pos_embed_for_scale).

Replaces, in ONE pass per stage, the reference sequence
  Projection (1x1 conv)            projection_head.py:134-149
  reshape + advanced-index gather  model_wo_extra_params.py:252-280 / model.py:300-306
  cat + permute                    model_wo_extra_params.py:281     / model.py:307-309
  scale-token concat               model_wo_extra_params.py:296-299 / model.py:322
  x + pos_embed_for_scale          scale_attention.py:331 / multi_vision_transformer.py:142-144
with a tcgen05 GEMM over the channels-last stage feature map whose epilogue adds bias and the
scale position embedding and scatters each pixel row straight to its (patch, scale) token row.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import engine, ops, trunk_convs
from .index_tables import num_scale_tokens, stages_used, token_row_maps


def _fold_batchnorm_(trunk: nn.Module) -> None:
    """Fold every eval-mode BatchNorm2d of a torchvision ResNet trunk into the convolution that
    feeds it (stem conv1/bn1, each Bottleneck's conv{1,2,3}/bn{1,2,3} and downsample.{0,1});
    the BN modules become Identity.  Works for both the index-named Sequential wrapper
    (children '0','1',...) and the name-based ResNetTrunkByScale."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval

    def fold_pairs(parent: nn.Module, pairs):
        for conv_name, bn_name in pairs:
            conv, bn = getattr(parent, conv_name, None), getattr(parent, bn_name, None)
            if isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d):
                setattr(parent, conv_name, fuse_conv_bn_eval(conv, bn))
                setattr(parent, bn_name, nn.Identity())

    if isinstance(trunk, nn.Sequential):
        fold_pairs(trunk, [("0", "1")])
    else:
        fold_pairs(trunk, [("conv1", "bn1")])
    for m in trunk.modules():
        if m.__class__.__name__ in ("Bottleneck", "BasicBlock"):
            fold_pairs(m, [("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3")])
            ds = getattr(m, "downsample", None)
            if isinstance(ds, nn.Sequential) and len(ds) == 2:
                fold_pairs(ds, [("0", "1")])


def _conv_args(conv: nn.Conv2d):
    return list(conv.stride), list(conv.padding), list(conv.dilation), conv.groups


def _fused_bottleneck(blk: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """BN-folded torchvision Bottleneck through cuDNN's fused conv+bias+ReLU / conv+add+bias+ReLU
    (ATen cudnn_convolution_relu / cudnn_convolution_add_relu): no separate bias / ReLU / add kernels."""
    out = torch.cudnn_convolution_relu(x, blk.conv1.weight, blk.conv1.bias, *_conv_args(blk.conv1))
    out = torch.cudnn_convolution_relu(out, blk.conv2.weight, blk.conv2.bias, *_conv_args(blk.conv2))
    bias3 = blk.conv3.bias
    ds = blk.downsample
    if ds is None:
        identity = x
    elif isinstance(ds, nn.Sequential) and len(ds) == 2 and isinstance(ds[0], nn.Conv2d) and ds[0].bias is not None \
            and isinstance(ds[1], nn.Identity):
        # BN-folded projection shortcut: its bias joins conv3's (relu(conv3(.) + b3 + conv_ds(x) + b_ds)), so the
        # shortcut is a bias-free convolution and no elementwise add kernel runs (0.85 ms per 256 images)
        conv = ds[0]
        merged = getattr(blk, "_duo_merged_bias", None)
        if merged is None or merged.device != bias3.device or merged.dtype != bias3.dtype:
            merged = (bias3.float() + conv.bias.float()).to(bias3.dtype)
            blk._duo_merged_bias = merged
        identity = torch.nn.functional.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        bias3 = merged
    else:
        identity = ds(x)
    return torch.cudnn_convolution_add_relu(out, blk.conv3.weight, identity, 1.0, bias3, *_conv_args(blk.conv3))


def _stem_pool(pool: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """The stem's MaxPool2d(3, 2, 1): own NHWC kernel (torch's channels-last max-pool runs at ~10 % of the
    HBM roofline: 0.75 ms per 256 images); any other pooling configuration goes through the module."""
    def pair(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    std = (isinstance(pool, nn.MaxPool2d) and pair(pool.kernel_size) == (3, 3) and pair(pool.stride) == (2, 2)
           and pair(pool.padding) == (1, 1) and pair(pool.dilation) == (1, 1) and not pool.ceil_mode)
    if (std and x.is_cuda and x.dtype in (torch.float16, torch.bfloat16) and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last)):
        return ops.maxpool3x3s2(x)
    return pool(x)


def _fused_trunk_forward(t: nn.Module, x: torch.Tensor, by_scale: bool) -> Dict[int, torch.Tensor]:
    """Forward of a BN-folded ResNet-50 trunk with fused cuDNN epilogues; returns the four stage maps."""
    if by_scale:
        stem, pool = t.conv1, t.maxpool
        layers = [t.layer1, t.layer2, t.layer3, t.layer4]
    else:
        ch = dict(t.named_children())
        stem, pool = ch["0"], ch["3"]
        layers = [ch["4"], ch["5"], ch["6"], ch["7"]]
    x = torch.cudnn_convolution_relu(x, stem.weight, stem.bias, *_conv_args(stem))
    x = _stem_pool(pool, x)
    feats: Dict[int, torch.Tensor] = {}
    for i, layer in enumerate(layers):
        for blk in layer:
            x = _fused_bottleneck(blk, x)
        feats[i] = x
    return feats


# fp16 range guard of the trunk (bf16 mode).  The BN-folded trunk is positively homogeneous once its biases are
# scaled with the input (convolution, ReLU, max-pool and the residual add all commute with a positive factor), so
# running it on alpha * x with alpha * bias yields alpha * (every activation) EXACTLY for a power of two alpha.
# alpha is chosen once per weight set from the largest activation magnitude of a bf16 calibration pass (bf16 has
# fp32's range) so that it lands near 2^10 — a factor 64 below the fp16 maximum 65504 — and 1 / alpha is folded into
# the 1x1 projection weights of the token builder.  alpha == 1 (every activation <= 2^12, the usual case) leaves the
# path bit-identical to an unguarded one.
_FP16_SAFE_MAX = 4096.0   # no rescaling below this calibration maximum
_FP16_TARGET_MAX = 1024.0  # rescaled activations peak near this value


class TrunkRunner:
    """Runs the ResNet trunk in the precision of the path (fp16 channels-last BN-folded copy of the fp32 master
    weights, re-made when they change) and returns the tapped stage maps, scaled by `act_scale`.

    bf16 mode, ResNet-50-style trunks: every convolution runs on the package's own implicit-GEMM kernel
    (trunk_convs.OwnTrunk, csrc/conv_tcgen05.cu); `backend = "cudnn"` selects the fused cuDNN calls instead (other
    trunks — the r18 BasicBlocks — always use them).  fp32 mode stays on cuDNN fp32 modules."""

    def __init__(self):
        self._sig = None
        self._trunk: Optional[nn.Module] = None
        self._own: Optional[trunk_convs.OwnTrunk] = None
        self.backend = "own"    # "own": conv_tcgen05 kernels where the trunk is eligible; "cudnn": fused cuDNN calls
        self._verified = False  # fast path checked against the plain module path for this weight set
        # dtype of the cuDNN trunk in bf16 mode: "fp16" (default: same speed as bf16, 11-bit mantissa —
        # the ~50 stacked convolutions otherwise contribute ~1e-2 of the 2e-2 bf16 error budget before
        # the first transformer block), "bf16", or "fp32"
        self.trunk_dtype = "fp16"
        # power-of-two factor the stage maps are scaled by (fp16 trunk only; 1.0 otherwise); set by calibration on
        # the first forward of a weight set, or pinned by the caller through `act_scale_override`
        self.act_scale = 1.0
        self.act_scale_override: Optional[float] = None
        self.calibration_max: Optional[float] = None

    def _dtype(self, precision: str) -> torch.dtype:
        if precision == "fp32":
            return torch.float32
        return {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[self.trunk_dtype]

    @staticmethod
    def _calibrate(folded_fp32: nn.Module, x: torch.Tensor, by_scale: bool) -> float:
        """Largest |activation| any convolution of the BN-folded trunk produces for (a sample of) x, measured in bf16."""
        t = copy.deepcopy(folded_fp32).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        peaks: List[torch.Tensor] = []
        hooks = [m.register_forward_hook(lambda _m, _i, o: peaks.append(o.detach().abs().amax().float()))
                 for m in t.modules() if isinstance(m, nn.Conv2d)]
        xs = x[:16].to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        TrunkRunner._plain_forward(t, xs, by_scale)
        for h in hooks:
            h.remove()
        return float(torch.stack(peaks).max().item())

    def _packed_trunk(self, trunk: nn.Module, precision: str, x: torch.Tensor, by_scale: bool) -> nn.Module:
        dt = self._dtype(precision)
        sig = engine.param_signature(trunk, precision + str(dt) + str(self.act_scale_override) + self.backend)
        if self._trunk is None or self._sig != sig:
            t = copy.deepcopy(trunk).eval().float()
            scale = 1.0
            if dt != torch.float32:
                _fold_batchnorm_(t)  # eval-mode BN -> scale/shift of the preceding conv (fp32, before the cast)
            if dt == torch.float16:
                if self.act_scale_override is not None:
                    scale = float(self.act_scale_override)
                else:
                    self.calibration_max = self._calibrate(t, x, by_scale)
                    if not (self.calibration_max < float("inf")):
                        raise RuntimeError("trunk calibration: non-finite activations — check the input / checkpoint")
                    if self.calibration_max > _FP16_SAFE_MAX:
                        import math

                        scale = 2.0 ** -math.ceil(math.log2(self.calibration_max / _FP16_TARGET_MAX))
                if scale != 1.0:
                    for m in t.modules():
                        if isinstance(m, nn.Conv2d) and m.bias is not None:
                            m.bias.data.mul_(scale)
            own = None
            if dt != torch.float32 and self.backend == "own" and trunk_convs.eligible(t, by_scale):
                own = trunk_convs.OwnTrunk(t, by_scale, dt)  # packed from the fp32 folded weights: fp32 biases
            t = t.to(dtype=dt, memory_format=torch.channels_last)
            for p in t.parameters():
                p.requires_grad_(False)
            self._trunk, self._own, self._sig, self.act_scale, self._verified = t, own, sig, scale, False
        return self._trunk

    @torch.no_grad()
    def features(self, trunk: nn.Module, x: torch.Tensor, precision: str, by_scale: bool) -> Dict[int, torch.Tensor]:
        """Stage maps 0..3, each multiplied by `self.act_scale` (1.0 unless the fp16 range guard engaged)."""
        t = self._packed_trunk(trunk, precision, x, by_scale)
        dt = self._dtype(precision)
        x_in = x
        if self._own is None or not self._verified:
            if self.act_scale != 1.0:
                x = x * self.act_scale
            x = x.to(dtype=dt).contiguous(memory_format=torch.channels_last)
        old_tf32 = torch.backends.cudnn.allow_tf32
        if precision == "fp32":
            torch.backends.cudnn.allow_tf32 = False
        try:
            if dt == torch.float32:
                return self._plain_forward(t, x, by_scale)
            # Own implicit-GEMM convolutions, or BN-folded bottlenecks through cuDNN's fused conv+bias(+add)+ReLU.
            # Checked ONCE per weight set against the plain module path; a mismatch (or any error of the fast path) is
            # raised, never papered over.
            if self._own is not None:
                feats = self._own.features(x_in, self.act_scale)
            else:
                feats = _fused_trunk_forward(t, x, by_scale)
            if not self._verified:
                ref = self._plain_forward(t, x, by_scale)
                for k in ref:
                    a, r = feats[k].float(), ref[k].float()
                    if not torch.isfinite(a).all():
                        raise RuntimeError(f"trunk stage {k}: non-finite {dt} activations (act_scale={self.act_scale}, "
                                           f"calibration max {self.calibration_max}) — pin TrunkRunner.act_scale_override")
                    err = float((a - r).abs().max() / r.abs().max().clamp_min(1e-30))
                    if err > 1e-2:
                        raise RuntimeError(f"trunk stage {k}: {'own convolution' if self._own is not None else 'fused cuDNN'} path "
                                           f"differs from the module path by {err:.2e}")
                self._verified = True
            return feats
        finally:
            torch.backends.cudnn.allow_tf32 = old_tf32

    @staticmethod
    def _plain_forward(t: nn.Module, x: torch.Tensor, by_scale: bool) -> Dict[int, torch.Tensor]:
        if by_scale:  # ResNetTrunkByScale returns [layer1..layer4]  (resnet50ssl.py:35-45)
            outs = t(x)
            return {i: o for i, o in enumerate(outs)}
        feats: Dict[int, torch.Tensor] = {}
        # nn.Sequential(conv1,bn1,relu,maxpool,layer1..4): children '4'..'7' are the stage taps
        # (model_wo_extra_params.py:214-224, model.py:213-223)
        for name, module in t.named_children():
            x = module(x)
            if name in ("4", "5", "6", "7"):
                feats[int(name) - 4] = x
        return feats


def _unscaled(feats: Dict[int, torch.Tensor], act_scale: float) -> Dict[int, torch.Tensor]:
    """Stage maps with the trunk's fp16 range-guard factor removed (bf16: fp32's range), for consumers other than
    the token-builder GEMM (the channel-token branch).  The usual act_scale == 1 returns the maps untouched."""
    if act_scale == 1.0:
        return feats
    return {k: (v.float() * (1.0 / act_scale)).to(torch.bfloat16) for k, v in feats.items()}


class TokenBuilder(engine.PackCache):
    def __init__(self):
        self._maps: Dict[Tuple[int, int, str], Dict[int, torch.Tensor]] = {}

    def row_maps(self, num_layers: int, g: int, device: torch.device) -> Dict[int, torch.Tensor]:
        key = (num_layers, g, str(device))
        if key not in self._maps:
            self._maps[key] = {k: m.to(device) for k, m in token_row_maps(num_layers, g).items()}
        return self._maps[key]

    def pack(self, projection: nn.Module, num_layers: int, precision: str,
             dtype: torch.dtype = torch.bfloat16, act_scale: float = 1.0) -> Dict[int, Tuple]:
        """1x1 projection weights as GEMM operands; 1 / act_scale (a power of two: exact) folded in when the stage
        maps arrive scaled by the trunk's fp16 range guard."""
        def build():
            return {k: engine.pack_linear(projection.head(k).weight.detach().float() / act_scale, projection.head(k).bias,
                                          precision, dtype)
                    for k in stages_used(num_layers)}

        return self.packed(build, projection, precision + str(dtype) + str(act_scale))

    @torch.no_grad()
    def build(
        self,
        feats: Dict[int, torch.Tensor],
        projection: nn.Module,
        num_layers: int,
        scale_tok: torch.Tensor,
        pos_scale: torch.Tensor,
        precision: str,
        act_scale: float = 1.0,
    ) -> torch.Tensor:
        """feats[k]: [B, C_k, g*w, g*w] (any memory format; channels-last is free), multiplied by act_scale.
        scale_tok: [D] learned channel_token, or [B, P, D] channel-branch output (fp32).
        pos_scale: [S, D] fp32.  Returns X fp32 [B, P, S, D]."""
        B, _, h3, w3 = feats[3].shape
        assert h3 == w3, "square inputs only"
        g = h3
        P = g * g
        S = num_scale_tokens(num_layers)
        D = pos_scale.shape[1]
        dev = feats[3].device
        X = torch.empty(B, P, S, D, dtype=torch.float32, device=dev)
        ops.fill_scale_token(X, scale_tok, pos_scale[0])
        maps = self.row_maps(num_layers, g, dev)
        # fp16 trunk maps feed the GEMM as they are, against fp16 copies of the 1x1-conv weights
        op_dtype = torch.float16 if (precision == "bf16" and feats[3].dtype == torch.float16) else torch.bfloat16
        packs = self.pack(projection, num_layers, precision, op_dtype, act_scale)
        for k in stages_used(num_layers):
            f = feats[k]
            Bk, C, H, W = f.shape
            assert H == g * 2 ** (3 - k) and W == H, f"stage {k}: expected {g * 2 ** (3 - k)}^2, got {H}x{W}"
            rows = f.permute(0, 2, 3, 1)  # NHWC view; contiguous when f is channels-last
            if precision == "bf16":
                A = rows.to(op_dtype).contiguous().view(Bk * H * W, C)
            else:
                a32 = rows.to(torch.float32).contiguous().view(Bk * H * W, C)
                A = torch.empty(Bk * H * W, 2 * C, dtype=torch.bfloat16, device=dev)
                ops.convert(a32, A)
            w, b = packs[k]
            ops.gemm(A, w, b, X.view(B * P * S, D), ops.EPI_SCATTER_F32, split3=(precision == "fp32"),
                     row_map=maps[k], rows_per_group=H * W, dest_rows_per_group=P * S, pos=pos_scale, pos_period=S)
        return X
