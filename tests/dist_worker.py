"""Worker of tests/test_distributed_gpu.py (launched by torchrun, one rank per GPU): the `wo4_d2` golden model on a
global batch sharded over the ranks through parallel.ShardedDuoFormer; every rank writes the gathered logits."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from common import build_product, load_golden  # noqa: E402
from duoformer_tcga_b200 import parallel  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    out_dir, batch = sys.argv[1], int(sys.argv[2])
    rank, local_rank, world = parallel.init_distributed("nccl", timeout_s=120)
    torch.cuda.set_device(local_rank)
    gold = load_golden("wo4_d2")
    model = build_product(gold["case"])
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=gold["weight_seed"]))
    model = model.cuda().eval()
    x = synth.synth_images(batch, seed=777).cuda()
    y = parallel.ShardedDuoFormer(model)(x)
    assert y.shape == (batch, 10) and y.is_cuda
    # host-pipeline path too: every rank feeds its own shard, receives the gathered logits
    lo, hi = parallel.shard_bounds(batch, rank, world)
    piped = list(parallel.HostPipeline(model, ragged=(batch % world != 0)).run([x[lo:hi].cpu().pin_memory()]))
    torch.save({"y": y.float().cpu(), "piped": piped[0], "device": torch.cuda.current_device()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
