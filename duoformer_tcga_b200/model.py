"""MyModel — DuoFormer "with extra params": channel-token branch + timm-derived
MultiscaleTransformer.  Drop-in for models/model.py:22-341 of the reference (same constructor
signature and state_dict keys).

Reference defects handled as documented in SURVEY.md App. A: D6 (only stages 2,3 are projected
in the reference, so only num_layers=2 runs there; here every stage num_layers needs is
projected), D7 (pos_embed sized from the real token count).  `img_size` is an additive keyword.
"""
from __future__ import annotations

import torch
import torchvision.models as models
from torch import nn

from . import engine
from .multi_vision_transformer import MultiscaleTransformer
from .projection_head import (Channel_Projector_All, Channel_Projector_layer1, Channel_Projector_layer2,
                              Channel_Projector_layer3, Projection)
from .scale_attention import _check_eval
from .channel_branch import ChannelBranch
from .token_builder import TokenBuilder, TrunkRunner, _unscaled


class MyModel(nn.Module):
    def __init__(
        self,
        depth=None,
        patch_size=49,
        embed_dim=256,
        num_heads=6,
        init_values=1e-5,
        num_classes=2,
        num_layers=4,
        proj_dim=512,
        model_ver="originalViT",
        pretrained=True,
        freeze=True,
        img_size=224,
    ):
        super().__init__()
        if embed_dim != proj_dim:
            raise ValueError(f"embed_dim ({embed_dim}) must equal proj_dim ({proj_dim})")
        if depth is None:
            raise TypeError("depth must be an int")
        self.name = model_ver
        self.num_layers = num_layers
        self.proj_dim = proj_dim
        if pretrained:
            trunk = models.resnet50(weights=models.ResNet50_Weights.DEFAULT)
        else:
            trunk = models.resnet50()
        self.resnet_projector = nn.Sequential(*list(trunk.children())[:-2])
        self.chann_proj1 = Channel_Projector_layer1()
        self.chann_proj2 = Channel_Projector_layer2()
        self.chann_proj3 = Channel_Projector_layer3()
        self.chann_proj_all = Channel_Projector_All()
        self.projection = Projection(num_layers=self.num_layers, proj_dim=self.proj_dim, backbone="r50")
        if self.num_layers > 1:
            self.vision_transformer = MultiscaleTransformer(
                pretrained=pretrained, depth=depth, scales=num_layers, num_heads=num_heads, patch_size=patch_size,
                embed_dim=embed_dim, init_values=init_values, num_classes=num_classes, model_type=self.name,
                attn_drop_rate=0.1, drop_rate=0.1, num_patches=(img_size // 32) ** 2,
            )
        else:
            raise NotImplementedError("MyModel needs num_layers > 1 (the reference builds no transformer otherwise)")
        if freeze:
            for param in self.resnet_projector.parameters():
                param.requires_grad = False
        self._trunk_runner = TrunkRunner()
        self._token_builder = TokenBuilder()
        self._channel_branch = None

    @property
    def precision(self) -> str:
        return self.vision_transformer.precision

    def set_precision(self, precision: str) -> "MyModel":
        if precision not in engine.PRECISIONS:
            raise ValueError(f"precision must be one of {engine.PRECISIONS}")
        self.vision_transformer.precision = precision
        return self

    @torch.no_grad()
    def get_features(self, x):
        """Stage feature maps keyed '0'..'3' (model_wo_extra_params.py:214-224 / model.py:213-223), unscaled."""
        tr = self._trunk_runner
        f = tr.features(self.resnet_projector, x, self.precision, False)
        if tr.act_scale != 1.0:  # the fp16 range guard engaged: hand back true magnitudes
            f = {k: v.float() / tr.act_scale for k, v in f.items()}
        return {str(k): v for k, v in f.items()}

    @torch.no_grad()
    def channel_branch(self, feats) -> torch.Tensor:
        """[B, P, D] fp32 channel token (model.py:279-289).  bf16 mode: implicit-GEMM convolutions on
        tcgen05 (channel_branch.py); fp32 mode: the fp32 cuDNN modules with TF32 off."""
        if self.precision == "bf16":
            if self._channel_branch is None:
                self._channel_branch = ChannelBranch(self.chann_proj1, self.chann_proj2, self.chann_proj_all)
            return self._channel_branch(feats)
        dt = torch.float32
        old = torch.backends.cudnn.allow_tf32
        if self.precision == "fp32":
            torch.backends.cudnn.allow_tf32 = False
        try:
            c0 = self.chann_proj1(feats[0].to(dt))
            c1 = self.chann_proj2(feats[1].to(dt))
            c2 = self.chann_proj3(feats[2].to(dt))
            fused = torch.cat([c0, c1, c2, feats[3].to(dt)], dim=1)
            tok = self.chann_proj_all(fused)
        finally:
            torch.backends.cudnn.allow_tf32 = old
        return tok.permute(0, 2, 1).contiguous()

    @torch.no_grad()
    def build_tokens(self, x: torch.Tensor) -> torch.Tensor:
        if self.name != "scaleformer":
            raise NotImplementedError("only model_ver='scaleformer' is on the DuoFormer path")
        tr = self._trunk_runner
        with engine.nvtx("trunk"):
            feats = tr.features(self.resnet_projector, x, self.precision, False)
        with engine.nvtx("channel_branch"):
            tok = self.channel_branch(_unscaled(feats, tr.act_scale))
        with engine.nvtx("token_builder"):
            return self._token_builder.build(feats, self.projection, self.num_layers, tok,
                                             self.vision_transformer.pos_scale_table(), self.precision, tr.act_scale)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _check_eval(self)
        engine.require_cuda(x, "MyModel.forward")
        if x.shape[0] == 0:  # empty batch: empty logits, no launches
            return torch.empty(0, self.vision_transformer.head.out_features, dtype=torch.float32, device=x.device)
        X = self.build_tokens(x)
        return self.vision_transformer.forward_prepared(X)


def count_parameters(model):
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    total = sum(p.numel() for p in model.parameters())
    return trainable / 1000000, total / 1000000
