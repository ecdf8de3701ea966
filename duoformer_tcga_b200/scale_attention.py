"""Scale attention + from-scratch multiscale transformer (wo-extra-params path).

Mirrors the classes, constructor signatures and state_dict keys of the reference's
models/scale_attention.py (AttentionForScale :23-45, ScaleBlock :48-93, AttentionForPatch
:176-211, PatchBlock :214-236, MultiscaleFormer :239-344); the forward math is executed by the
sm_100a kernels behind the C ABI (engine.py / ops.py).  Eval-mode semantics only.
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Dict, Optional

import torch
from torch import nn

from . import engine, ops
from .index_tables import num_scale_tokens
from .vit_layout import AttentionParams, LayerScale, Mlp, named_apply_vit_init, trunc_normal_


def _check_eval(module: nn.Module) -> None:
    if module.training:
        raise NotImplementedError(
            "duoformer_tcga_b200 implements the inference (eval-mode) forward only: call model.eval()"
        )


class AttentionForScale(AttentionParams):
    """Multi-head attention over the S scale tokens inside each patch (scale_attention.py:23-45)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0, proj_drop=0):
        # positional call of the reference: attn_drop lands in qk_norm (see AttentionParams)
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        if attn_drop:
            raise NotImplementedError("attn_drop_rate > 0 is not supported (SURVEY.md App. A D10)")


class ScaleBlock(nn.Module, engine.PackCache):
    """Pre-LN transformer block over the scale axis (scale_attention.py:48-93)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_norm=False, proj_drop=0.0,
                 attn_drop=0.0, init_values=None, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 mlp_layer=Mlp):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = AttentionForScale(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop,
                                      proj_drop=proj_drop)
        self.ls1 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), drop=proj_drop)
        self.ls2 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = nn.Identity()
        self.num_heads = num_heads
        self.precision = "bf16"

    def pack(self, precision: str) -> Dict:
        def build():
            at, mlp = self.attn, self.mlp
            return engine.pack_scale_block(
                precision, self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias,
                (at.qkv.weight, at.qkv.bias), (at.proj.weight, at.proj.bias),
                (mlp.fc1.weight, mlp.fc1.bias), (mlp.fc2.weight, mlp.fc2.bias),
                self.ls1.gamma if isinstance(self.ls1, LayerScale) else None,
                self.ls2.gamma if isinstance(self.ls2, LayerScale) else None)

        return self.packed(build, self, precision)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _check_eval(self)
        engine.require_cuda(x, "ScaleBlock.forward")
        X = x.to(torch.float32).contiguous().clone()
        ws = engine.Workspace(X.device)
        return engine.scale_stage(X, [self.pack(self.precision)], self.num_heads, self.attn.scale,
                                  self.norm1.eps, self.precision, ws, live_only_last=False)


class AttentionForPatch(AttentionParams):
    """Multi-head attention over the P+1 patch tokens (scale_attention.py:176-211)."""

    def __init__(self, dim=768, num_heads=8, qkv_bias=False, attn_drop=0, proj_drop=0):
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        if attn_drop:
            raise NotImplementedError("attn_drop_rate > 0 is not supported (SURVEY.md App. A D10)")


class PatchBlock(nn.Module, engine.PackCache):
    """Attention only — no LayerScale, MLP or residual (scale_attention.py:214-236)."""

    def __init__(self, dim, num_heads, qkv_bias=False, proj_drop=0.0, attn_drop=0.0):
        super().__init__()
        self.attn = AttentionForPatch(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop,
                                      proj_drop=proj_drop)

    def pack(self, precision: str) -> Dict:
        def build():
            pl = partial(engine.pack_linear, precision=precision)
            return {"qkv": pl(self.attn.qkv.weight, self.attn.qkv.bias),
                    "proj": pl(self.attn.proj.weight, self.attn.proj.bias)}

        return self.packed(build, self, precision)


class MultiscaleFormer(nn.Module):
    """depth scale blocks -> depth patch blocks -> head (scale_attention.py:239-344).

    `scale_token` / `patch_attn` are accepted and ignored: the reference passes them
    (model_wo_extra_params.py:104-105) although its class lacks them (App. A D4).  `fea_dim` for
    scales == 2 is the token count actually produced, 6 (the reference's 21 cannot broadcast, D5).
    """

    def __init__(
        self,
        depth: int = 12,
        scales: int = 2,
        num_heads: int = 12,
        embed_dim: int = 768,
        mlp_ratio: float = 4.0,
        qkv_bias: bool = True,
        qk_norm: bool = False,
        proj_drop_rate: float = 0.0,
        attn_drop_rate: float = 0.0,
        norm_layer: Optional[Callable] = None,
        act_layer: Optional[Callable] = None,
        init_values: Optional[float] = None,
        num_classes: int = 100,
        num_patches: int = 49,
        pos_drop_rate: float = 0.0,
        patch_drop_rate: float = 0.0,
        block_fn: Callable = ScaleBlock,
        block_fn1: Callable = PatchBlock,
        scale_token: str = "random",
        patch_attn: bool = True,
    ):
        super().__init__()
        self.norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.act_layer = act_layer or nn.GELU
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.num_patches = num_patches
        self.scaleBlocks = nn.Sequential(*[
            block_fn(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_norm=qk_norm,
                     init_values=init_values, proj_drop=proj_drop_rate, attn_drop=attn_drop_rate,
                     norm_layer=self.norm_layer, act_layer=self.act_layer)
            for _ in range(depth)
        ])
        self.blocks = nn.Sequential(*[
            block_fn1(dim=embed_dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_drop=proj_drop_rate,
                      attn_drop=attn_drop_rate)
            for _ in range(depth)
        ])
        self.fea_dim = num_scale_tokens(scales)  # 6 / 22 / 86
        embed_len = num_patches + 1
        self.pos_embed_for_scale = nn.Parameter(torch.randn(1, 1, self.fea_dim, embed_dim))
        self.pos_drop_for_scale = nn.Dropout(p=pos_drop_rate)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, embed_len, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(p=pos_drop_rate)
        self.fc_norm = self.norm_layer(embed_dim)  # dead in the reference forward (:342), kept for the schema
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(self.embed_dim, num_classes)
        self._init_weights()
        self.precision = "bf16"
        self.attn_algo = 0
        # skip work the reference computes but never consumes (last scale block: only s = 0 rows)
        self.dead_work_elimination = True
        # precision of the 12 residual-free patch blocks (< 1 % of the FLOPs): None = same as
        # `precision`; "fp32" runs them as 3-pass split GEMMs so that their bf16 rounding does not
        # compound through the stack (no residual stream to absorb it).
        self.patch_precision: Optional[str] = "fp32"
        # PatchBlock is proj(attention(qkv(x))) with NO residual / norm / activation (scale_attention.py:234-236), so block
        # i's proj and block i + 1's qkv are two adjacent linear maps: they are composed once per weight set
        # (W = W_qkv W_proj in fp64, b = W_qkv b_proj + b_qkv) and the 11 intermediate proj GEMMs of the patch stage
        # (and the re-rounding of their outputs) disappear.  Split-precision patch stage only (the default).  Nothing for
        # the 4-scale step (the patch stage is 1.5 % of it), 5 % for the 2-scale model where it is a fifth of the forward.
        self.fuse_patch_linears = True
        self._patch_cache = engine.PackCache()
        self._capture: Optional[Dict[str, torch.Tensor]] = None
        self._ws: Optional[engine.Workspace] = None

    def _init_weights(self):
        trunc_normal_(self.pos_embed_for_scale, std=0.036)
        trunc_normal_(self.pos_embed, std=0.036)
        nn.init.normal_(self.cls_token, std=0.036)
        named_apply_vit_init(self.blocks)
        named_apply_vit_init(self.scaleBlocks)

    # ---- engine entry points ------------------------------------------------------------
    def workspace(self, device: torch.device) -> engine.Workspace:
        if self._ws is None or self._ws.device != device:
            self._ws = engine.Workspace(device)
        return self._ws

    def pos_scale_table(self) -> torch.Tensor:
        """[S, D] fp32 view of pos_embed_for_scale (added by the token-builder epilogue)."""
        return self.pos_embed_for_scale.detach().reshape(self.fea_dim, self.embed_dim).to(torch.float32).contiguous()

    @torch.no_grad()
    def forward_prepared(self, X: torch.Tensor) -> torch.Tensor:
        """Forward from a token tensor that already contains `+ pos_embed_for_scale`
        (fp32 [B, P, S, D], modified in place)."""
        _check_eval(self)
        B, P, S, D = X.shape
        assert S == self.fea_dim and D == self.embed_dim and P + 1 == self.pos_embed.shape[1], (X.shape, self.fea_dim)
        prec = self.precision
        cap = self._capture
        if cap is not None:
            cap["tokens"] = X.clone()
        ws = self.workspace(X.device)
        scale = self.scaleBlocks[0].attn.scale if len(self.scaleBlocks) else 0.125
        engine.scale_stage(X, [b.pack(prec) for b in self.scaleBlocks], self.num_heads, scale,
                           self.scaleBlocks[0].norm1.eps if len(self.scaleBlocks) else 1e-6, prec, ws, cap,
                           attn_algo=self.attn_algo, live_only_last=self.dead_work_elimination)
        with engine.nvtx("patch_stage"):
            return self._patch_stage(X, ws, cap)

    def _patch_stage(self, X: torch.Tensor, ws: engine.Workspace, cap) -> torch.Tensor:
        B, P, S, D = X.shape
        prec = self.precision
        # patch stage: CLS + first scale token of every patch + pos_embed (scale_attention.py:183-193)
        N = P + 1
        prec = self.patch_precision or prec
        kd = 2 if prec in ("fp32", "mixed") else 1
        scratch = engine.PatchScratch(ws, B * N, D)  # the scale stage's workspace is free again
        Z = scratch.next_z((B * N, kd * D), torch.bfloat16)
        ops.assemble_patch_tokens(X, engine._f32(self.cls_token).view(-1), engine._f32(self.pos_embed).view(N, D), Z.view(B, N, kd * D))
        if cap is not None:
            cap["patch_in"] = engine.unsplit(Z, prec).view(B, N, D)
        nblk = len(self.blocks)
        cls_only = False
        fuse = self.fuse_patch_linears and prec == "fp32" and nblk >= 2
        packs = self._patch_packs(fuse) if fuse else None
        for i, blk in enumerate(self.blocks):
            last = i == nblk - 1
            cls_only = last and self.dead_work_elimination
            pk = packs[i] if fuse else blk.pack("bf16" if prec == "bf16" else "fp32")
            Z = engine.region_attention(Z, pk, N, self.num_heads, blk.attn.scale, prec, out_f32=last,
                                        cls_only=cls_only, scratch=scratch, skip_proj=fuse and not last)
            if cap is not None:
                if cls_only:
                    cap[f"patch_block_{i}_cls"] = Z.view(B, D).clone()
                elif fuse and not last:
                    # the block output x_i = proj_i(attention output) is not part of the fused data flow: computed for the
                    # probe only, from the attention output the next block's composed weights consume
                    xi = torch.empty(B * N, D, dtype=torch.float32, device=X.device)
                    ops.gemm(Z, pk["proj"][0], pk["proj"][1], xi, ops.EPI_F32, split3=1)
                    cap[f"patch_block_{i}"] = xi.view(B, N, D)
                else:
                    cap[f"patch_block_{i}"] = engine.unsplit(Z, prec).view(B, N, D).clone()
        if nblk == 0:
            Zf = engine.unsplit(Z, prec).contiguous()
        else:
            Zf = Z
        logits = torch.empty(B, self.head.out_features, dtype=torch.float32, device=X.device)
        # head on the CLS row; fc_norm is computed-and-discarded in the reference (:341-344)
        ops.head(Zf, D if cls_only else N * D, engine._f32(self.head.weight), engine._f32(self.head.bias), logits)
        return logits

    def _patch_packs(self, fuse: bool):
        """Split-precision operands of the patch blocks; with `fuse`, block i >= 1 carries qkv_i composed with proj_{i-1}."""
        def build():
            out = []
            for i, blk in enumerate(self.blocks):
                Wq = blk.attn.qkv.weight.detach().double()
                bq = blk.attn.qkv.bias.detach().double() if blk.attn.qkv.bias is not None else torch.zeros(Wq.shape[0], dtype=torch.float64, device=Wq.device)
                if fuse and i > 0:
                    prev = self.blocks[i - 1].attn.proj
                    Wp = prev.weight.detach().double()
                    if prev.bias is not None:
                        bq = Wq @ prev.bias.detach().double() + bq
                    Wq = Wq @ Wp
                out.append({"qkv": engine.pack_linear(Wq.float(), bq.float(), "fp32"),
                            "proj": engine.pack_linear(blk.attn.proj.weight, blk.attn.proj.bias, "fp32")})
            return out

        return self._patch_cache.packed(build, self.blocks, f"fp32-fuse{int(fuse)}")

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference entry point: x [B, P, S, D] tokens WITHOUT the scale position embedding."""
        _check_eval(self)
        engine.require_cuda(x, "MultiscaleFormer.forward")
        x = x.to(torch.float32).contiguous()
        X = torch.empty_like(x)
        ops.add_pos(x, self.pos_scale_table(), X)
        return self.forward_prepared(X)
