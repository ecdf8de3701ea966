"""Informational (-m gpu, opt-in with DUO_RUN_LIBRARY_BAR=1): the "library bar" of BASELINE.md §5 — the
reference's algorithm (the oracle's functional forward: plain torch ops over cuBLAS / cuDNN) run eagerly
on the same B200 in bf16 and fp32, next to this repo's path, on the bench workload (4-scale, depth 12).
Writes gpurun_out/library_bar.json; asserts only that the CUDA path is not slower than eager bf16."""
import json
import os

import pytest
import torch

from common import COMMON
import duoformer_tcga_b200 as duo
from oracle import duoformer_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(os.environ.get("DUO_RUN_LIBRARY_BAR") != "1", reason="opt-in measurement")
def test_library_bar():
    B = int(os.environ.get("DUO_LIBRARY_BAR_BATCH", "64"))
    torch.manual_seed(0)
    model = duo.MyModel_no_extra_params(depth=12, num_layers=4, pretrained=False, **COMMON).eval()
    sd32 = {k: v.cuda() for k, v in orc.cpu_state_dict(model).items()}
    sd16 = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v) for k, v in sd32.items()}
    x = torch.randn(B, 3, 224, 224, device="cuda")
    model = model.cuda()

    def timeit(fn, iters=3):
        with torch.no_grad():
            fn(); fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    res = {"batch": B}
    res["this_repo_bf16_ms"] = timeit(lambda: model(x))
    res["torch_eager_bf16_ms"] = timeit(lambda: orc.forward_wo_extra(x.to(torch.bfloat16), sd16, 12, 12, 4))
    res["torch_eager_fp32_tf32_ms"] = timeit(lambda: orc.forward_wo_extra(x, sd32, 12, 12, 4))
    for k in list(res):
        if k.endswith("_ms"):
            res[k.replace("_ms", "_images_per_s")] = round(B / res[k] * 1000, 1)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/library_bar.json", "w"), indent=1)
    print(json.dumps(res))
    assert res["this_repo_bf16_ms"] < res["torch_eager_bf16_ms"]
