"""Times the other BASELINE.json configs on one B200 (informational; bench.py is the headline).

  cfg3  wo-extra 2-scale (S=6), batch 128, bf16 and fp32 mode
  cfg4  wo-extra 4-scale at 384x384 (P=144), batch 128, bf16
  mm2   MyModel 2-scale (channel-token branch on cuDNN), batch 128, bf16
  cfg5  wo-extra 4-scale, batch 1024 on ONE GPU (chunked token workspace)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import duoformer_tcga_b200 as duo  # noqa: E402
from duoformer_tcga_b200 import ops  # noqa: E402

COMMON = dict(embed_dim=768, num_heads=12, num_classes=10, proj_dim=768)


def timeit(model, x, iters=3, warmup=2):
    with torch.no_grad():
        for _ in range(warmup):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            y = model(x)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, y


def main():
    out = []
    torch.manual_seed(0)
    cases = [
        ("cfg3_2scale_b128_bf16", dict(kind="wo", num_layers=2, num_patches=49), 128, 224, "bf16"),
        ("cfg3_2scale_b128_fp32mode", dict(kind="wo", num_layers=2, num_patches=49), 128, 224, "fp32"),
        ("cfg4_4scale_384_b128_bf16", dict(kind="wo", num_layers=4, num_patches=144), 128, 384, "bf16"),
        ("mymodel_2scale_b128_bf16", dict(kind="mm", num_layers=2), 128, 224, "bf16"),
        ("cfg5_4scale_b1024_one_gpu_bf16", dict(kind="wo", num_layers=4, num_patches=49), 1024, 224, "bf16"),
    ]
    only = sys.argv[1:]
    for name, cfg, B, size, prec in cases:
        if only and not any(o in name for o in only):
            continue
        if cfg["kind"] == "wo":
            m = duo.MyModel_no_extra_params(depth=12, num_layers=cfg["num_layers"], num_patches=cfg["num_patches"],
                                            pretrained=False, **COMMON)
        else:
            m = duo.MyModel(depth=12, patch_size=32, num_layers=2, model_ver="scaleformer", pretrained=False, **COMMON)
        m = m.cuda().eval().set_precision(prec)
        x = torch.randn(B, 3, size, size, device="cuda")
        ms, y = timeit(m, x)
        # one profiled forward: CUDA events around every GEMM / LayerNorm / attention launch of the library
        prof = []
        ops.PROFILE = prof
        with torch.no_grad():
            m(x)
        ops.PROFILE = None
        torch.cuda.synchronize()
        by_tag = {}
        for a, b, kind, fl, nb, tag in prof:
            t = by_tag.setdefault(f"{kind}:{tag}", [0.0, 0])
            t[0] += a.elapsed_time(b)
            t[1] += 1
        top = sorted(by_tag.items(), key=lambda kv: -kv[1][0])[:8]
        r = {"config": name, "batch": B, "ms_per_forward": round(ms, 3), "images_per_s": round(B / ms * 1000, 1),
             "finite": bool(torch.isfinite(y).all()), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2),
             "library_ms_by_launch_kind": {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in top},
             "library_ms_total": round(sum(v[0] for v in by_tag.values()), 3)}
        print(json.dumps(r), flush=True)
        out.append(r)
        del m, x, y
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/bench_configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
