"""CPU (-m "not gpu"): the N>1 host path — batch sharding and the rank-major logit all-gather —
with world_size 2 over gloo (the same code runs over NCCL on the GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import ROOT  # noqa: F401
from duoformer_tcga_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _ToyModel(torch.nn.Module):
    """Stands in for the per-rank forward: logits are a deterministic function of each image."""

    def forward(self, x):
        return torch.stack([x.flatten(1).sum(1) * (c + 1) for c in range(10)], dim=1)


def _worker(rank, world, port, batch, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    parallel.init_distributed("gloo")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 3, 4, 4, generator=g)
    model = parallel.ShardedDuoFormer(_ToyModel())
    y = model(x)
    torch.save(y, os.path.join(out_dir, f"y{rank}.pt"))
    lo, hi = parallel.shard_bounds(batch, rank, world)
    assert parallel.shard_batch(x, rank, world).shape[0] == hi - lo
    # shard sizes known only to their owners (HostPipeline(ragged=True)): exchanged, padded, trimmed
    y2 = parallel.all_gather_logits(_ToyModel()(x[lo:hi]), ragged=True)
    assert torch.equal(y2, y)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_sharded_forward_equals_single_process(tmp_path, batch):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 3, 4, 4, generator=g)
    ref = _ToyModel()(x)
    for r in range(world):
        y = torch.load(os.path.join(str(tmp_path), f"y{r}.pt"))
        assert torch.equal(y, ref), f"rank {r}: gathered logits differ from the single-process result"


def test_shard_bounds_cover_batch_exactly():
    for batch in (1, 7, 8, 2048):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
