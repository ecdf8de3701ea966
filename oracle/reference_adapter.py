"""Import the UNMODIFIED reference (/root/reference/models) under the timm stand-in —
TEST INFRASTRUCTURE ONLY, build container only (the directory does not exist on the GPU box).

Patches applied from outside, never editing the reference:
  D1  `models/` put on sys.path as a str so the flat imports resolve;
  D3  torchvision resnet50/resnet18 constructors replaced by weights=None versions (no network);
  D4  MultiscaleFormer sub-class that swallows scale_token / patch_attn;
  D5  for scales == 2 pos_embed_for_scale re-created with 6 rows.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import torch
from torch import nn

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


_mods = None


def load_reference():
    """Returns the reference's flat modules (model, model_wo_extra_params, scale_attention, ...)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("/root/reference is not present")
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "timm_shim")
    for p in (shim, os.path.join(REFERENCE_ROOT, "models"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    saved_env = os.environ.get("CUDA_VISIBLE_DEVICES")  # backbone.py:13 sets it as an import side effect (D16)
    with contextlib.redirect_stdout(io.StringIO()):
        import model as ref_model  # noqa
        import model_wo_extra_params as ref_wo  # noqa
        import multi_vision_transformer as ref_mvt  # noqa
        import projection_head as ref_ph  # noqa
        import scale_attention as ref_sa  # noqa
    if saved_env is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = saved_env

    import torchvision.models as tvm

    class _Models:  # D3
        @staticmethod
        def resnet50(*a, **k):
            return tvm.resnet50(weights=None)

        @staticmethod
        def resnet18(*a, **k):
            return tvm.resnet18(weights=None)

    ref_wo.models = _Models

    base = ref_sa.MultiscaleFormer

    class PatchedFormer(base):  # D4 + D5
        def __init__(self, *a, scale_token="random", patch_attn=True, **k):
            super().__init__(*a, **k)
            if k.get("scales") == 2:
                self.fea_dim = 6
                self.pos_embed_for_scale = nn.Parameter(torch.randn(1, 1, 6, self.embed_dim))

    ref_wo.MultiscaleFormer = PatchedFormer
    _mods = {"model": ref_model, "wo": ref_wo, "mvt": ref_mvt, "ph": ref_ph, "sa": ref_sa}
    return _mods


def build_wo_extra(**kw):
    m = load_reference()["wo"]
    with contextlib.redirect_stdout(io.StringIO()):
        return m.MyModel_no_extra_params(**kw).eval()


def build_mymodel(**kw):
    m = load_reference()["model"]
    with contextlib.redirect_stdout(io.StringIO()):
        return m.MyModel(**kw).eval()
