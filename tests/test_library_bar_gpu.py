"""-m gpu: the "library bar" of BASELINE.md §5 — the same forward as plain torch ops (cuDNN / cuBLAS / SDPA, eager,
tools/library_bar.py) on the same B200 and parameters, next to this repo's path.  bench.py reports it as
`library_bar`; here: the eager forward computes the same function, and the CUDA path is not slower than it."""
import json
import os

import pytest
import torch

from common import build_product, load_golden, relerr
from oracle import synth
from tools import library_bar

pytestmark = pytest.mark.gpu


def test_library_bar_same_function_and_slower_than_this_repo():
    gold = load_golden("wo4_d2")
    model = build_product(gold["case"])
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=gold["weight_seed"]))
    model = model.cuda().eval()
    B = 16
    res, eager = library_bar.measure(model, batch=B, iters=2)
    x = synth.synth_images(B, seed=5).cuda()
    with torch.no_grad():
        y = model(x).float()
        assert relerr(eager(x), y) < 4e-2  # two bf16 paths of the same function
        e32 = library_bar.EagerDuoFormer(model, torch.float32)
        assert relerr(y[:4], e32(x[:4])) < 2e-2
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        model(x)
        e0.record()
        for _ in range(2):
            model(x)
        e1.record()
        torch.cuda.synchronize()
    ours = B / (e0.elapsed_time(e1) / 2) * 1000.0
    res["this_repo_images_per_s"] = ours
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/library_bar_test.json", "w"), indent=1)
    assert res["finite"] and ours > res["value"], res
