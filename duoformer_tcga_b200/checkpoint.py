"""Checkpoint ingestion for reference-trained DuoFormer weights.

The reference saves `torch.save({"epoch", "model": <whole nn.Module>, "optimizer", ...})`
(main_toy.py:139-149) — a pickle of the module object, which can only be un-pickled where the
classes are importable at their original paths (`model.MyModel`, `scale_attention.MultiscaleFormer`,
`timm.models.vision_transformer.Block`, ...).  `load_checkpoint` accepts

  * a plain state_dict (or a path to one),
  * a dict holding it under "model" / "state_dict" / "model_state_dict",
  * the reference's whole-module pickle — un-pickled through stand-in classes (any attribute of the
    reference's / timm's module paths resolves to an empty nn.Module subclass, which is all that
    `state_dict()` needs), so neither the reference nor timm has to be installed,

normalises the keys (DataParallel "module." prefix; index-based vs name-based ResNet trunk naming,
App. B) and tolerates the reference's dead keys.  Released checkpoints are not reachable offline:
this path is verified by round-tripping a reference-built module in the build container
(tests/test_checkpoint.py), not against a published file — "checkpoint parity unpinned".
"""
from __future__ import annotations

import contextlib
import importlib.machinery
import re
import sys
import types
from typing import Dict, Iterable, Tuple, Union

import torch
from torch import nn

# keys that exist in some reference checkpoints but never influence the forward (App. A D10/D12, App. B)
DEAD_KEY_PATTERNS = (
    r"\.num_batches_tracked$",
    r"^vision_transformer\.fc_norm\.",
    r"^vision_transformer\.patch_embed\.proj\.",
    r"\.attn\.(q_norm|k_norm)\.",
)

_TRUNK_INDEX_TO_NAME = {"0": "conv1", "1": "bn1", "4": "layer1", "5": "layer2", "6": "layer3", "7": "layer4"}
_TRUNK_NAME_TO_INDEX = {v: k for k, v in _TRUNK_INDEX_TO_NAME.items()}

# module paths the reference's pickles point into (flat imports with models/ on sys.path, and the
# package spelling), plus timm's
_STUB_ROOTS = ("model", "model_wo_extra_params", "scale_attention", "multiscale_attn", "multi_vision_transformer",
               "projection_head", "resnet50ssl", "backbone", "models", "timm")


class _StubModule(types.ModuleType):
    """A module whose every attribute is an (empty) nn.Module subclass of that name, and whose
    sub-modules are stubs too — enough for pickle to rebuild module objects."""

    def __init__(self, name: str):
        super().__init__(name)
        self.__path__ = []  # behave like a package
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, item: str):
        if item.startswith("__"):
            raise AttributeError(item)
        cls = type(item, (nn.Module,), {"__module__": self.__name__, "forward": lambda self, *a, **k: None})
        setattr(self, item, cls)
        return cls


class _StubFinder:
    """meta_path finder: any import below one of the stub roots yields a _StubModule."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        return None


@contextlib.contextmanager
def reference_unpickle_stubs():
    """Temporarily make the reference's and timm's module paths importable as stand-ins (only for
    roots that are not genuinely importable already)."""
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _STUB_ROOTS}
    finder = _StubFinder()
    sys.meta_path.append(finder)  # after the real finders: a genuinely installed timm wins
    try:
        yield
    finally:
        sys.meta_path.remove(finder)
        for k in [k for k, v in sys.modules.items() if isinstance(v, _StubModule)]:
            del sys.modules[k]
        sys.modules.update(saved)


def extract_state_dict(obj) -> Dict[str, torch.Tensor]:
    """state_dict from any of the accepted containers."""
    if isinstance(obj, nn.Module):
        return obj.state_dict()
    if isinstance(obj, dict):
        for key in ("model", "state_dict", "model_state_dict"):
            if key in obj and isinstance(obj[key], (dict, nn.Module)):
                return extract_state_dict(obj[key])
        if obj and all(isinstance(v, torch.Tensor) for v in obj.values()):
            return obj
    raise TypeError(f"cannot find a state_dict in an object of type {type(obj).__name__}")


def _convert_trunk_key(key: str, want_named: bool) -> str:
    m = re.match(r"^resnet_projector\.([^.]+)\.(.*)$", key)
    if not m:
        return key
    head, rest = m.group(1), m.group(2)
    if want_named and head in _TRUNK_INDEX_TO_NAME:
        return f"resnet_projector.{_TRUNK_INDEX_TO_NAME[head]}.{rest}"
    if not want_named and head in _TRUNK_NAME_TO_INDEX:
        return f"resnet_projector.{_TRUNK_NAME_TO_INDEX[head]}.{rest}"
    return key


def normalise_keys(sd: Dict[str, torch.Tensor], model: nn.Module) -> Dict[str, torch.Tensor]:
    """Strip DataParallel prefixes and convert the trunk naming to what `model` uses."""
    target = model.state_dict().keys()
    want_named = any(k.startswith("resnet_projector.conv1.") for k in target)
    out = {}
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[len("module."):]
        out[_convert_trunk_key(k, want_named)] = v
    return out


def is_dead_key(key: str) -> bool:
    return any(re.search(p, key) for p in DEAD_KEY_PATTERNS)


def load_checkpoint(model: nn.Module, source: Union[str, dict, nn.Module], strict: bool = True,
                    trust_pickle: bool = False) -> Tuple[Iterable[str], Iterable[str]]:
    """Load reference weights into a duoformer_tcga_b200 model.

    A path is first read with torch.load(weights_only=True) (state_dict files: tensors only, no code execution).
    Whole-module pickles — what the reference's training script saves (main_toy.py:170-183) — execute arbitrary
    pickle code on load and are only read when the caller passes trust_pickle=True.

    Returns (missing, unexpected) AFTER discounting the reference's dead keys; with strict=True a
    RuntimeError is raised if either list is non-empty or a shape differs."""
    if isinstance(source, str):
        path = source
        try:
            source = torch.load(path, map_location="cpu", weights_only=True)
        except Exception as e:  # not a plain tensor container
            if not trust_pickle:
                raise RuntimeError(f"{path} is not a tensors-only checkpoint ({type(e).__name__}); a whole-module pickle "
                                   "runs arbitrary code when loaded — pass trust_pickle=True if you trust its origin") from e
            with reference_unpickle_stubs():
                source = torch.load(path, map_location="cpu", weights_only=False)
    sd = normalise_keys(extract_state_dict(source), model)
    own = model.state_dict()
    bad_shape = [k for k, v in sd.items() if k in own and tuple(own[k].shape) != tuple(v.shape)]
    if bad_shape:
        raise RuntimeError("shape mismatch for: " + ", ".join(
            f"{k} {tuple(sd[k].shape)} vs {tuple(own[k].shape)}" for k in bad_shape[:8]))
    result = model.load_state_dict(sd, strict=False)
    missing = [k for k in result.missing_keys if not is_dead_key(k)]
    unexpected = [k for k in result.unexpected_keys if not is_dead_key(k)]
    if strict and (missing or unexpected):
        raise RuntimeError(f"checkpoint does not match the model: missing {missing[:8]} unexpected {unexpected[:8]}")
    return missing, unexpected
