"""Shared helpers for the test-suite (oracle = checker only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import duoformer_tcga_b200 as duo  # noqa: E402
from oracle import duoformer_oracle as orc  # noqa: E402
from oracle import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
COMMON = dict(embed_dim=768, num_heads=12, num_classes=10, proj_dim=768)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def build_product(case, **extra):
    """Product model (CPU-constructed, random init) for a golden-case config."""
    if case["kind"] == "wo":
        return duo.MyModel_no_extra_params(depth=case["depth"], num_layers=case["num_layers"],
                                           backbone=case["backbone"], scale_token=case["scale_token"],
                                           pretrained=False, **COMMON, **extra).eval()
    return duo.MyModel(depth=case["depth"], patch_size=32, init_values=1e-5, num_layers=case["num_layers"],
                       model_ver="scaleformer", pretrained=False, **COMMON, **extra).eval()


def oracle_forward(case, x, sd, capture=None):
    if case["kind"] == "wo":
        return orc.forward_wo_extra(x, sd, case["depth"], COMMON["num_heads"], case["num_layers"],
                                    backbone=case["backbone"], scale_token=case["scale_token"], capture=capture)
    return orc.forward_mymodel(x, sd, case["depth"], COMMON["num_heads"], case["num_layers"], capture=capture)


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def probe_values(t, probe):
    return t.detach().float().cpu().reshape(-1)[probe["idx"]]
