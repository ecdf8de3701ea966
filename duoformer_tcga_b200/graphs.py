"""CUDA-graph replay of the whole forward for launch-bound (small-batch) serving.

At batch 2 a DuoFormer forward is ~130 kernel launches of a few microseconds each: the GPU waits
for Python.  Every launch of this package is capture-safe (enqueue-only, no host syncs, TMA
descriptors passed by value), so the forward — the trunk's convolution launches included — can be captured once per
input shape and replayed with a single `cudaGraphLaunch`.

A captured graph bakes in device ADDRESSES.  Everything it reads or writes is therefore owned (or pinned) by the
GraphedForward object: a private activation workspace (the model's own one is swapped out during warm-up and
capture and restored afterwards, so later eager forwards of any size cannot free or regrow it), references to the
packed weights / trunk copy / row maps that existed at capture, and the parameter + precision signature — a replay
after the weights or the precision changed raises instead of running stale operands.
"""
from __future__ import annotations

from typing import List

import torch

from . import engine


def _transformers(model: torch.nn.Module) -> List[torch.nn.Module]:
    return [m for m in model.modules() if hasattr(m, "_ws") and hasattr(m, "workspace")]


class GraphedForward:
    """Captures `model(x)` for one input shape; `__call__` copies the input into the static
    buffer, replays the graph and returns a copy of the logits."""

    def __init__(self, model: torch.nn.Module, example_input: torch.Tensor, warmup: int = 3):
        if not example_input.is_cuda:
            raise NotImplementedError("GraphedForward needs a CUDA input (no CPU fallback)")
        self.model = model
        self.static_in = example_input.clone()
        dev = example_input.device
        owners = _transformers(model)
        saved = [m._ws for m in owners]
        self._workspaces = []
        for m in owners:  # private workspace: the graph's activation buffers never alias the eager path's
            m._ws = engine.Workspace(dev)
            self._workspaces.append(m._ws)
        try:
            with torch.cuda.device(dev):
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side), torch.no_grad():
                    for _ in range(warmup):  # packs weights, sets kernel attributes, calibrates / verifies the trunk path
                        model(self.static_in)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph), torch.no_grad():
                    self.static_out = model(self.static_in)
        finally:
            for m, ws in zip(owners, saved):
                m._ws = ws
        # pin what the graph addresses: workspace buffers, packed operands, the trunk copy, token row maps
        self._pins = [ws.buf for ws in self._workspaces]
        for m in list(model.modules()) + [getattr(model, n, None) for n in ("_trunk_runner", "_token_builder", "_channel_branch")]:
            for attr in ("_packed", "_trunk", "_own", "_maps", "_patch_cache"):
                v = getattr(m, attr, None)
                if v is not None:
                    self._pins.append(getattr(v, "_packed", v) if attr == "_patch_cache" else v)
        self._signature = self._model_signature()

    def _model_signature(self):
        vt = getattr(self.model, "vision_transformer", self.model)
        return engine.param_signature(self.model, f"{getattr(vt, 'precision', '')}|{getattr(vt, 'patch_precision', '')}|"
                                                  f"{getattr(vt, 'dead_work_elimination', '')}|{getattr(vt, 'fuse_patch_linears', '')}|{engine.FORWARD_LN_STATS}")

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape != self.static_in.shape:
            raise ValueError(f"captured for input shape {tuple(self.static_in.shape)}, got {tuple(x.shape)}")
        if self._model_signature() != self._signature:
            raise RuntimeError("the model's parameters or precision changed since the graph was captured: "
                               "capture a new GraphedForward")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out.clone()
