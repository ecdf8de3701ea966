"""CUDA-graph replay of the whole forward for launch-bound (small-batch) serving.

At batch 2 a DuoFormer forward is ~130 kernel launches of a few microseconds each: the GPU waits
for Python.  Every launch of this package is capture-safe (enqueue-only, no host syncs, TMA
descriptors passed by value), so the forward — cuDNN trunk included — can be captured once per
input shape and replayed with a single `cudaGraphLaunch`.
"""
from __future__ import annotations

import torch


class GraphedForward:
    """Captures `model(x)` for one input shape; `__call__` copies the input into the static
    buffer, replays the graph and returns a copy of the logits."""

    def __init__(self, model: torch.nn.Module, example_input: torch.Tensor, warmup: int = 3):
        if not example_input.is_cuda:
            raise NotImplementedError("GraphedForward needs a CUDA input (no CPU fallback)")
        self.model = model
        self.static_in = example_input.clone()
        side = torch.cuda.Stream(device=example_input.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # packs weights, sets kernel attributes, verifies the fused trunk path
                model(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = model(self.static_in)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape != self.static_in.shape:
            raise ValueError(f"captured for input shape {tuple(self.static_in.shape)}, got {tuple(x.shape)}")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out.clone()
