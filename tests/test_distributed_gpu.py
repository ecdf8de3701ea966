"""-m gpu, needs >= 2 GPUs (skipped otherwise): the data-parallel path over NCCL.  SURVEY.md §4 item 5 — 1/2/4/8-GPU
runs must produce the same gathered logits as one GPU, rank-major."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from common import ROOT, build_product, load_golden, relerr
from duoformer_tcga_b200 import parallel
from oracle import synth

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("batch", [8, 7])
def test_nccl_gathered_logits_equal_single_gpu_rank_major(tmp_path, batch):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if world < 4 else (4 if world < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path), str(batch)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    gold = load_golden("wo4_d2")
    model = build_product(gold["case"])
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=gold["weight_seed"]))
    model = model.cuda().eval()
    x = synth.synth_images(batch, seed=777).cuda()
    with torch.no_grad():
        # rank-major reference: the same shards, each run alone on this GPU (same kernels, same batch composition)
        shards = [model(parallel.shard_batch(x, rk, world)).float().cpu().reshape(-1, 10) for rk in range(world)
                  if parallel.shard_bounds(batch, rk, world)[1] > parallel.shard_bounds(batch, rk, world)[0]]
        ref = torch.cat(shards, dim=0)
        whole = model(x).float().cpu()
    devices = set()
    for rk in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"rank{rk}.pt"))
        devices.add(d["device"])
        assert torch.equal(d["y"], ref), f"rank {rk}: gathered logits differ from the single-GPU shards"
        assert torch.equal(d["piped"], ref), f"rank {rk}: HostPipeline gathered logits differ"
        assert relerr(d["y"], whole) < 5e-3  # whole-batch forward: cuDNN may differ by an fp16 ulp with batch position
    assert len(devices) == world


def test_model_on_second_device_while_first_is_current():
    """The library launches on the CURRENT device; ops switch to the operands' device first (a model on
    cuda:1 with cuda:0 current used to launch device-0 kernels on device-1 pointers)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    gold = load_golden("wo4_d2")
    model = build_product(gold["case"])
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=gold["weight_seed"]))
    x = synth.synth_images(gold["case"]["batch"], seed=gold["input_seed"])
    torch.cuda.set_device(0)
    with torch.no_grad():
        y0 = model.to("cuda:0").eval()(x.to("cuda:0")).float().cpu()
        assert torch.cuda.current_device() == 0
        y1 = model.to("cuda:1")(x.to("cuda:1"))
        assert y1.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(y1.float().cpu(), y0)
    assert relerr(y0, gold["logits"]) < 2e-2
    with pytest.raises(RuntimeError, match="different devices"):
        from duoformer_tcga_b200 import ops

        ops.layernorm(torch.zeros(4, 768, device="cuda:0"), torch.ones(768, device="cuda:1"), torch.zeros(768, device="cuda:0"),
                      torch.empty(4, 768, dtype=torch.bfloat16, device="cuda:0"), 1e-6)
