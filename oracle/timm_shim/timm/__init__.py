"""Minimal stand-in for timm==0.9.8 (environmental.yml:156) — TEST INFRASTRUCTURE ONLY.

timm is not installed in this image and cannot be (no network).  This package restates, from the
published timm 0.9.8 semantics, exactly the symbols the reference imports so
that /root/reference/models can be imported UNMODIFIED by oracle/make_golden.py to generate the
golden vectors that pin the oracle.  It is never imported by the product package.
"""
from . import layers, models  # noqa: F401


def create_model(*args, **kwargs):  # only used by the out-of-scope ViTBase16 baseline
    raise RuntimeError("timm.create_model is not available in the timm stand-in")
