from . import _manipulate, resnetv2, vision_transformer  # noqa: F401
