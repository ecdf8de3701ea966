"""MultiscaleTransformer (MyModel path): scale blocks, then region attention, norm, head.

Mirrors models/multi_vision_transformer.py:19-171 of the reference, which sub-classes timm's
`VisionTransformer` and therefore inherits (and keeps in its state_dict) an unused
`patch_embed` conv, `cls_token`, `pos_embed`, `norm` and `head`.  Forward semantics replicated:

  x + pos_embed_for_scale (:142-144); depth x forward_change_order_attn1 (:145-146);
  region attention of block 0 with CLS + pos_embed (:151-156); region attention of blocks
  1..depth-1 each applied to the SAME x, so only the last one reaches the output (:157-159,
  SURVEY.md App. A D13 — the dead ones are skipped); norm (:161); head(...).squeeze() (:169-171).
"""
from __future__ import annotations

from functools import partial
from typing import Dict, Optional

import torch
from torch import nn

from . import engine, ops
from .index_tables import num_scale_tokens
from .multiscale_attn import MultiscaleBlock
from .scale_attention import _check_eval
from .vit_layout import PatchEmbedParams, named_apply_vit_init, trunc_normal_


class MultiscaleTransformer(nn.Module):
    def __init__(
        self,
        pretrained=False,
        depth=12,
        scales=2,
        num_heads=6,
        patch_size=16,
        embed_dim=384,
        mlp_ratio=4.0,
        qkv_bias=True,
        qk_norm=False,
        drop_rate=0.0,
        drop_path_rate=0.0,
        attn_drop_rate=0.0,
        norm_layer=None,
        act_layer=None,
        init_values=1e-5,
        num_classes=1000,
        model_type="scaleformer",
        num_patches: Optional[int] = None,
    ):
        super().__init__()
        # ---- what timm VisionTransformer.__init__(depth, patch_size, num_classes, embed_dim, num_heads)
        #      leaves behind (App. B) ----
        self.patch_embed = PatchEmbedParams(img_size=224, patch_size=patch_size, in_chans=3, embed_dim=embed_dim)
        # App. A D7: the position table must cover the real token count P+1; timm would size it
        # from (224 // patch_size)**2 + 1, which is only right for patch_size == 32.
        self.num_patches = num_patches if num_patches is not None else self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.num_patches + 1, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(p=0.0)
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)

        self.dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.act_layer = act_layer or nn.GELU
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.model = model_type
        self.blocks = nn.Sequential(*[
            MultiscaleBlock(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                            qk_norm=qk_norm, init_values=init_values, proj_drop=drop_rate,
                            attn_drop=attn_drop_rate, drop_path=self.dpr[i], norm_layer=self.norm_layer,
                            act_layer=self.act_layer)
            for i in range(depth)
        ])
        self.fea_dim = num_scale_tokens(scales)
        self.cls_token1 = None
        self.pos_embed_for_scale = nn.Parameter(torch.randn(1, 1, self.fea_dim, embed_dim))
        self.pos_drop_for_scale = nn.Dropout(p=0.0)
        self._init_weights()
        self.precision = "bf16"
        self.attn_algo = 0
        # skip work the reference computes but never consumes (last scale block: only s = 0 rows)
        self.dead_work_elimination = True
        self._capture: Optional[Dict[str, torch.Tensor]] = None
        self._ws: Optional[engine.Workspace] = None

    def _init_weights(self):
        named_apply_vit_init(self)  # timm ViT init over every Linear (incl. head)
        trunc_normal_(self.pos_embed_for_scale, std=0.036)
        named_apply_vit_init(self.blocks)

    def workspace(self, device: torch.device) -> engine.Workspace:
        if self._ws is None or self._ws.device != device:
            self._ws = engine.Workspace(device)
        return self._ws

    def pos_scale_table(self) -> torch.Tensor:
        return self.pos_embed_for_scale.detach().reshape(self.fea_dim, self.embed_dim).to(torch.float32).contiguous()

    @torch.no_grad()
    def forward_prepared(self, X: torch.Tensor) -> torch.Tensor:
        """Forward from fp32 tokens [B,P,S,D] that already include pos_embed_for_scale (in place)."""
        _check_eval(self)
        if self.model != "scaleformer":
            raise NotImplementedError("only model_type='scaleformer' is on the DuoFormer path")
        B, P, S, D = X.shape
        assert S == self.fea_dim and D == self.embed_dim and P + 1 == self.pos_embed.shape[1], (X.shape, self.fea_dim)
        prec = self.precision
        cap = self._capture
        if cap is not None:
            cap["tokens"] = X.clone()
        depth = len(self.blocks)
        packs = [b.pack(prec) for b in self.blocks]
        scale = self.blocks[0].attn.scale if depth else 1.0
        engine.scale_stage(X, packs, self.num_heads, scale, self.blocks[0].norm1.eps if depth else 1e-6, prec,
                           self.workspace(X.device), cap, attn_algo=self.attn_algo, live_only_last=self.dead_work_elimination)
        N = P + 1
        kd = 2 if prec == "fp32" else 1
        logits = torch.empty(B, self.head.out_features, dtype=torch.float32, device=X.device)
        hw, hb = engine._f32(self.head.weight), engine._f32(self.head.bias)
        nw, nb = engine._f32(self.norm.weight), engine._f32(self.norm.bias)
        if depth >= 1:
            scratch = engine.PatchScratch(self.workspace(X.device), B * N, D)  # the scale stage's workspace is free again
            Z = scratch.next_z((B * N, kd * D), torch.bfloat16)
            ops.assemble_patch_tokens(X, engine._f32(self.cls_token).view(-1), engine._f32(self.pos_embed).view(N, D),
                                      Z.view(B, N, kd * D))
            Z = engine.region_attention(Z, packs[0]["region"], N, self.num_heads, scale, prec, out_f32=False, scratch=scratch)
            if cap is not None:
                cap["region_block_0"] = engine.unsplit(Z, prec).view(B, N, D).clone()
        if depth >= 2:
            Zf = engine.region_attention(Z, packs[-1]["region"], N, self.num_heads, scale, prec, out_f32=True, scratch=scratch)
            if cap is not None:
                cap["region_block_last"] = Zf.view(B, N, D).clone()
            ops.head(Zf, N * D, hw, hb, logits, ln_gamma=nw, ln_beta=nb, eps=self.norm.eps)
        else:
            # depth <= 1: the reference's cls_token variable is still the raw (expanded) parameter (:137-161)
            ops.head(engine._f32(self.cls_token).view(-1), 0, hw, hb, logits, ln_gamma=nw, ln_beta=nb, eps=self.norm.eps)
        return logits.squeeze()

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference entry point: x [B, P, S, D] tokens WITHOUT the scale position embedding."""
        _check_eval(self)
        engine.require_cuda(x, "MultiscaleTransformer.forward")
        x = x.to(torch.float32).contiguous()
        X = torch.empty_like(x)
        ops.add_pos(x, self.pos_scale_table(), X)
        return self.forward_prepared(X)
