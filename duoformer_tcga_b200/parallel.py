"""Batch-sharded data parallelism: one process per GPU, weights replicated, images sharded,
a single all-gather of the [B/G, num_classes] fp32 logits (rank-major = original order).

The reference has no distributed code at all; images are independent units on
this path (eval-mode BatchNorm, no cross-sample op), so there is no data-path collective other
than gathering the logits — NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (no-op for a single process)."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank `rank`; the first batch % world ranks get one extra image."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def all_gather_logits(local_logits: torch.Tensor, global_batch: Optional[int] = None) -> torch.Tensor:
    """[B_r, C] per rank -> [sum_r B_r, C] on every rank, rank-major.  Equal shards use one
    all_gather_into_tensor; ragged shards are padded to the largest shard and trimmed."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_logits
    world = dist.get_world_size()
    C = local_logits.shape[1]
    if global_batch is None or global_batch % world == 0:
        out = torch.empty(world * local_logits.shape[0], C, dtype=local_logits.dtype, device=local_logits.device)
        dist.all_gather_into_tensor(out, local_logits.contiguous())
        return out
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, C, dtype=local_logits.dtype, device=local_logits.device)
    pad[: local_logits.shape[0]] = local_logits
    out = torch.empty(world * mx, C, dtype=local_logits.dtype, device=local_logits.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx : r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


class ShardedDuoFormer(torch.nn.Module):
    """Wraps a DuoFormer model: forward(global_batch) runs this rank's slice and returns the
    gathered logits of the whole batch (identical on every rank)."""

    def __init__(self, model: torch.nn.Module):
        super().__init__()
        self.model = model

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return self.model(x)
        rank, world = dist.get_rank(), dist.get_world_size()
        local = shard_batch(x, rank, world)
        y = self.model(local)
        if y.dim() == 1:
            y = y.unsqueeze(0)
        return all_gather_logits(y, x.shape[0])
