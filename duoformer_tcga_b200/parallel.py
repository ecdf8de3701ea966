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


def init_distributed(backend: Optional[str] = None, timeout_s: Optional[float] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (no-op for a single process).
    timeout_s: collective timeout (tests use a short one so that a mismatched collective fails fast)."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        import datetime

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {} if timeout_s is None else {"timeout": datetime.timedelta(seconds=timeout_s)}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank), **kw)
        else:
            dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank `rank`; the first batch % world ranks get one extra image."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def all_gather_logits(local_logits: torch.Tensor, global_batch: Optional[int] = None,
                      ragged: bool = False) -> torch.Tensor:
    """[B_r, C] per rank -> [sum_r B_r, C] on every rank, rank-major.  Equal shards (the default assumption when
    global_batch is None) use one all_gather_into_tensor.  Ragged shards are padded to the largest shard and
    trimmed: their sizes follow from global_batch (shard_bounds), or — ragged=True, sizes only known to their
    owners — are exchanged first (one extra tiny collective and a host read of the sizes)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_logits
    world = dist.get_world_size()
    C = local_logits.shape[1]
    if ragged and global_batch is None:
        mine = torch.tensor([local_logits.shape[0]], dtype=torch.int64, device=local_logits.device)
        every = torch.empty(world, dtype=torch.int64, device=local_logits.device)
        dist.all_gather_into_tensor(every, mine)
        counts = every.tolist()
        sizes, lo = [], 0
        for n in counts:
            sizes.append((lo, lo + n))
            lo += n
    elif global_batch is None or global_batch % world == 0:
        out = torch.empty(world * local_logits.shape[0], C, dtype=local_logits.dtype, device=local_logits.device)
        dist.all_gather_into_tensor(out, local_logits.contiguous())
        return out
    else:
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


class HostPipeline:
    """Feeds host batches to a DuoFormer model and returns host logits: the host->device copy of batch
    i+1 runs on its own stream while batch i is computed (two device input buffers); the logits of every
    batch are copied to pinned host memory as soon as they exist and handed out one step late, so the
    launch queue of the GPU never drains between batches.  On several GPUs every rank feeds its own
    shard and receives the gathered logits of the whole step (equal shard sizes on all ranks unless ragged=True).

        pipe = HostPipeline(model)            # model: cuda, eval
        for logits in pipe.run(batches):      # batches: iterable of (ideally pinned) host tensors [b,3,H,W]
            ...                               # logits: host fp32 [b * world, num_classes]
    """

    def __init__(self, model: torch.nn.Module, device: Optional[torch.device] = None, gather: bool = True,
                 ragged: bool = False):
        self.model = model
        self.device = device if device is not None else next(model.parameters()).device
        assert self.device.type == "cuda", "HostPipeline feeds a CUDA model"
        self.gather = gather
        self.ragged = ragged  # ranks may feed shards of different sizes in one step (sizes exchanged per step)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._bufs = [None, None]
        self._consumed = [None, None]  # event: the forward that read buffer k has been enqueued and finished
        self._compute: Optional[torch.cuda.Stream] = None  # stream the forwards run on (set by run())

    def _stage(self, k: int, host: torch.Tensor):
        """Enqueue the copy of `host` into device buffer k on the copy stream; returns (tensor, ready event)."""
        buf = self._bufs[k]
        fresh = buf is None or buf.shape != host.shape or buf.dtype != host.dtype
        with torch.cuda.stream(self.copy_stream):
            if fresh:
                # A new device buffer may reuse a block the caching allocator just took back from the compute
                # stream (activations of the forward still in flight): allocate it on the copy stream's pool and let
                # the copy wait for everything enqueued on the compute stream so far.
                fence = torch.cuda.Event()
                fence.record(self._compute)
                self.copy_stream.wait_event(fence)
                buf = torch.empty(host.shape, dtype=host.dtype, device=self.device)
                self._bufs[k] = buf
            if self._consumed[k] is not None and not fresh:
                self.copy_stream.wait_event(self._consumed[k])  # the previous user of this buffer is done
            buf.copy_(host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        buf.record_stream(self._compute)  # read by kernels of the compute stream
        return buf, ready

    @torch.no_grad()
    def run(self, host_batches):
        it = iter(host_batches)
        first = next(it, None)
        if first is None:
            return
        k = 0
        compute = torch.cuda.current_stream(self.device)
        self._compute = compute
        staged = self._stage(k, first)
        pending = None  # (pinned host logits, event) of the previous batch
        while staged is not None:
            x, ready = staged
            nxt = next(it, None)
            staged_next = self._stage(k ^ 1, nxt) if nxt is not None else None  # overlaps the forward below
            compute.wait_event(ready)
            y = self.model(x)
            done = torch.cuda.Event()
            done.record(compute)
            self._consumed[k] = done
            if y.dim() == 1:
                y = y.unsqueeze(0)
            if self.gather and dist.is_initialized() and dist.get_world_size() > 1:
                y = all_gather_logits(y, ragged=self.ragged)
            y = y.float()
            y_host = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
            y_host.copy_(y, non_blocking=True)  # device -> host read of this step's result
            landed = torch.cuda.Event()
            landed.record(compute)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0]
            pending = (y_host, landed)
            staged, k = staged_next, k ^ 1
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]
