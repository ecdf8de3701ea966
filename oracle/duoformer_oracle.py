"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A self-contained fp32 PyTorch (CPU-capable) restatement of the reference's DuoFormer forward
(AliSerwat/duoformer_TCGA), written functionally over a state_dict with the reference's key
schema.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it — as the checker / CPU baseline, never as the thing shipped.

Why a restatement: part of the reference's arithmetic lives in timm==0.9.8
(environmental.yml:156), which is not installed and cannot be (no network).  The functions
below restate exactly the timm pieces the reference inherits (Attention/Block/Mlp/LayerScale
semantics) together with the reference's own modules, each citing the
reference file:line it follows.

Pinning: the reference has NO tests, golden vectors or fixtures.  The oracle
is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container under a
minimal timm stand-in (oracle/timm_shim) with the three documented patches (App. A D3/D4/D5):
oracle/make_golden.py generates tests/golden/*.pt from the real reference modules and
tests/test_oracle.py checks this file against them.  The 384x384 (g = 12) generalisation has no
reference semantics (hard-coded 7x7) and is "parity unpinned": the oracle is its definition.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------
# index tables — model_wo_extra_params.py:110-212 (int64) / model.py:106-211 (int32), g = 7
# literal tables; generalised to any patch grid g
# --------------------------------------------------------------------------------------------
def index_table(k: int, g: int = 7) -> torch.Tensor:
    w = 2 ** (3 - k)
    G = g * w
    rows = []
    for r in range(g):
        for c in range(g):
            if k == 2:  # :117-124 — TL, BL, TR, BR (column-major)
                offs = [(0, 0), (1, 0), (0, 1), (1, 1)]
            else:  # :125-212 — row-major
                offs = [(dr, dc) for dr in range(w) for dc in range(w)]
            rows.append([(w * r + dr) * G + (w * c + dc) for dr, dc in offs])
    return torch.tensor(rows, dtype=torch.int64)


def num_scale_tokens(num_layers: int) -> int:
    return 1 + sum(4**i for i in range(num_layers))


# --------------------------------------------------------------------------------------------
# trunk — torchvision ResNet children '0'..'7', taps '4'..'7' -> '0'..'3'
# model_wo_extra_params.py:214-224 ; resnet50ssl.py:35-45 (name-based keys for r50_Swav)
# --------------------------------------------------------------------------------------------
def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, eps)


def _bottleneck(x, sd, p, stride):
    idt = x
    o = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1."))
    o = F.relu(_bn(F.conv2d(o, sd[p + "conv2.weight"], stride=stride, padding=1), sd, p + "bn2."))
    o = _bn(F.conv2d(o, sd[p + "conv3.weight"]), sd, p + "bn3.")
    if p + "downsample.0.weight" in sd:
        idt = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1.")
    return F.relu(o + idt)


def trunk_features(x: torch.Tensor, sd: SD, prefix: str = "resnet_projector.", by_name: bool = False) -> Dict[str, torch.Tensor]:
    """ResNet-50 (torchvision v1.5 layout: stride on the 3x3 conv) stage maps '0'..'3'."""
    names = (["conv1", "bn1", "layer1", "layer2", "layer3", "layer4"] if by_name else ["0", "1", "4", "5", "6", "7"])
    c1, b1, l1, l2, l3, l4 = [prefix + n + "." for n in names]
    x = F.conv2d(x, sd[c1 + "weight"], stride=2, padding=3)
    x = F.relu(_bn(x, sd, b1))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = {}
    for si, (lp, nblk) in enumerate(zip([l1, l2, l3, l4], [3, 4, 6, 3])):
        for bi in range(nblk):
            stride = 2 if (bi == 0 and si > 0) else 1
            x = _bottleneck(x, sd, f"{lp}{bi}.", stride)
        feats[str(si)] = x
    return feats


# --------------------------------------------------------------------------------------------
# channel-token branch — projection_head.py:152-268 ; call sites model.py:279-289,
# model_wo_extra_params.py:236-248
# --------------------------------------------------------------------------------------------
def channel_token(feats: Dict[str, torch.Tensor], sd: SD) -> torch.Tensor:
    def conv(x, p, stride, pad):
        return F.conv2d(x, sd[p + "weight"], sd[p + "bias"], stride=stride, padding=pad)

    c0 = F.max_pool2d(conv(conv(feats["0"], "chann_proj1.conv1.", 2, 1), "chann_proj1.conv2.", 2, 1), 2, 2)  # :169-176
    c1 = F.max_pool2d(conv(feats["1"], "chann_proj2.conv1.", 2, 1), 2, 2)  # :198-203
    c2 = F.max_pool2d(feats["2"], 2, 2)  # :220-222
    x = torch.cat([c0, c1, c2, feats["3"]], dim=1)  # sorted keys '0'..'3' (model.py:284-286)
    for i in range(4):  # ConvBatchNorm x4 (:242-268)
        p = f"chann_proj_all.nConvs.{i}."
        x = F.relu(_bn(conv(x, p + "conv.", 1, 1), sd, p + "norm."))
    x = torch.flatten(x, start_dim=2)  # [B, D, P]
    return x.unsqueeze(-1).permute(0, 2, 3, 1)  # [B, P, 1, D]  (model.py:287-289)


# --------------------------------------------------------------------------------------------
# token builder — projection_head.py:134-149 + model_wo_extra_params.py:252-299 / model.py:291-322
# --------------------------------------------------------------------------------------------
def build_tokens(feats: Dict[str, torch.Tensor], sd: SD, num_layers: int, scale_tok: torch.Tensor, g: int = 7) -> torch.Tensor:
    """-> [B, P, S, D] WITHOUT pos_embed_for_scale (exactly what the reference hands to the ViT)."""
    B = feats["3"].shape[0]
    parts = []
    for k in [3, 2, 1, 0][:num_layers]:
        w, b = sd[f"projection.proj_heads{k}.weight"], sd[f"projection.proj_heads{k}.bias"]
        y = F.conv2d(feats[str(k)], w, b)  # 1x1 conv
        C = y.shape[1]
        y = y.reshape(B, C, -1)[:, :, index_table(k, g)]  # [B, C, P, w*w]
        parts.append(y)
    x = torch.cat(parts, dim=-1).permute(0, 2, 3, 1)  # [B, P, S-1, D]
    P = g * g
    if scale_tok.dim() == 4 and scale_tok.shape[0] == 1:
        scale_tok = scale_tok.expand(B, P, -1, -1)  # model_wo_extra_params.py:299
    return torch.cat((scale_tok, x), dim=2)


# --------------------------------------------------------------------------------------------
# transformer pieces
# --------------------------------------------------------------------------------------------
def _ln(x, sd, p, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps)


def _lin(x, sd, p):
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


def scale_attention(x, sd, qkv_p, proj_p, H, scale):
    """AttentionForScale.forward scale_attention.py:28-45 / forward_with_scale multiscale_attn.py:149-166."""
    B, P, S, C = x.shape
    qkv = _lin(x, sd, qkv_p).reshape(B, P, S, 3, H, C // H).permute(3, 0, 1, 4, 2, 5)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    x = (attn @ v).transpose(2, 3).reshape(B, P, S, C)
    return _lin(x, sd, proj_p)


def region_attention(z, sd, qkv_p, proj_p, H, scale):
    """AttentionForPatch.forward scale_attention.py:195-209 / forward_with_region multiscale_attn.py:205-219
    (N generalised from the hard-coded 50)."""
    B, N, C = z.shape
    qkv = _lin(z, sd, qkv_p).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    z = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    return _lin(z, sd, proj_p)


def _mlp(x, sd, p):
    """timm Mlp: fc2(GELU_erf(fc1(x))), dropouts are identity in eval."""
    return _lin(F.gelu(_lin(x, sd, p + "fc1.")), sd, p + "fc2.")


def multiscale_former(tokens: torch.Tensor, sd: SD, depth: int, H: int, prefix: str = "vision_transformer.",
                      capture: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """MultiscaleFormer.forward scale_attention.py:330-344 (wo-extra path).  Softmax scale =
    head_dim**-0.5 inherited from timm Attention."""
    p = prefix
    C = tokens.shape[-1]
    scale = (C // H) ** -0.5
    x = tokens + sd[p + "pos_embed_for_scale"]  # :331
    if capture is not None:
        capture["tokens"] = x.clone()
    for i in range(depth):  # ScaleBlock.forward :90-93 (ls = Identity: init_values=None)
        b = f"{p}scaleBlocks.{i}."
        x = x + scale_attention(_ln(x, sd, b + "norm1."), sd, b + "attn.qkv.", b + "attn.proj.", H, scale)
        x = x + _mlp(_ln(x, sd, b + "norm2."), sd, b + "mlp.")
        if capture is not None:
            capture[f"scale_block_{i}"] = x.clone()
    B = x.shape[0]
    cls = sd[p + "cls_token"].expand(B, -1, -1)  # :335 (+ squeeze(1) at :190)
    z = torch.cat((cls, x[:, :, 0, :]), dim=1) + sd[p + "pos_embed"]  # :183-193
    if capture is not None:
        capture["patch_in"] = z.clone()
    for i in range(depth):  # PatchBlock: attention only :234-236
        b = f"{p}blocks.{i}."
        z = region_attention(z, sd, b + "attn.qkv.", b + "attn.proj.", H, scale)
        if capture is not None:
            capture[f"patch_block_{i}"] = z.clone()
    cls = z[:, 0, :]  # :341 ; fc_norm computed then discarded :342
    return _lin(cls, sd, p + "head.")  # :343-344


def multiscale_transformer(tokens: torch.Tensor, sd: SD, depth: int, H: int, prefix: str = "vision_transformer.",
                           capture: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """MultiscaleTransformer.forward multi_vision_transformer.py:133-171 (MyModel path).  Softmax
    scale 2*dim**-0.5 (multiscale_attn.py:142); LayerScale gammas when present."""
    p = prefix
    C = tokens.shape[-1]
    scale = 2 * C**-0.5
    B = tokens.shape[0]
    x = tokens + sd[p + "pos_embed_for_scale"]  # :142-144
    if capture is not None:
        capture["tokens"] = x.clone()
    for i in range(depth):  # forward_change_order_attn1 multiscale_attn.py:282-285
        b = f"{p}blocks.{i}."
        a = scale_attention(_ln(x, sd, b + "norm1."), sd, b + "attn.qkv1.", b + "attn.proj1.", H, scale)
        x = x + (a * sd[b + "ls1.gamma"] if b + "ls1.gamma" in sd else a)
        m = _mlp(_ln(x, sd, b + "norm2."), sd, b + "mlp.")
        x = x + (m * sd[b + "ls2.gamma"] if b + "ls2.gamma" in sd else m)
        if capture is not None:
            capture[f"scale_block_{i}"] = x.clone()
    cls_token = sd[p + "cls_token"].expand(B, -1, -1, -1)  # [B,1,1,C] :137-139
    z = None
    for i in range(depth):
        b = f"{p}blocks.{i}."
        if i == 0:  # forward_change_order_attn2_block1 :287-289 with CLS + pos_embed
            z = torch.cat((cls_token.squeeze(1), x[:, :, 0, :]), dim=1) + sd[p + "pos_embed"]
            z = region_attention(z, sd, b + "attn.qkv.", b + "attn.proj.", H, scale)
            if capture is not None:
                capture["region_block_0"] = z.clone()
        else:  # forward_change_order_attn2 :291-295 — z is NOT updated (:158)
            r = region_attention(z, sd, b + "attn.qkv.", b + "attn.proj.", H, scale)
            cls_token = r[:, 0, :]
            if capture is not None and i == depth - 1:
                capture["region_block_last"] = r.clone()
    cls_token = _ln(cls_token, sd, p + "norm.")  # :161
    return _lin(cls_token, sd, p + "head.").squeeze()  # :169-171


# --------------------------------------------------------------------------------------------
# whole models
# --------------------------------------------------------------------------------------------
def forward_wo_extra(x: torch.Tensor, sd: SD, depth: int, num_heads: int, num_layers: int,
                     backbone: str = "r50", scale_token: str = "random",
                     capture: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """MyModel_no_extra_params.forward model_wo_extra_params.py:226-302."""
    g = x.shape[-1] // 32
    feats = trunk_features(x, sd, by_name=(backbone == "r50_Swav"))
    if scale_token == "channel":
        tok = channel_token(feats, sd)
    else:
        tok = sd["channel_token"]
    tokens = build_tokens(feats, sd, num_layers, tok, g)
    if capture is not None:
        capture["features"] = feats
    return multiscale_former(tokens, sd, depth, num_heads, capture=capture)


def forward_mymodel(x: torch.Tensor, sd: SD, depth: int, num_heads: int, num_layers: int = 2,
                    capture: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """MyModel.forward model.py:225-341 (model_ver='scaleformer')."""
    g = x.shape[-1] // 32
    feats = trunk_features(x, sd)
    tok = channel_token(feats, sd)
    tokens = build_tokens(feats, sd, num_layers, tok, g)
    if capture is not None:
        capture["features"] = feats
    return multiscale_transformer(tokens, sd, depth, num_heads, capture=capture)


def cpu_state_dict(model_or_sd) -> SD:
    sd = model_or_sd if isinstance(model_or_sd, dict) else model_or_sd.state_dict()
    return {k: v.detach().to("cpu", torch.float32) if v.is_floating_point() else v.detach().cpu() for k, v in sd.items()}
