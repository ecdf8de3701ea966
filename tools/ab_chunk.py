"""Chunk-size sweep of the scale stage (engine.SCALE_CHUNK_TOKENS): forward time of the bench model (batch 256) as a
CUDA graph (the small-chunk settings issue ~10^4 launches per forward: eager would measure Python) and eagerly."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import duoformer_tcga_b200 as duo  # noqa: E402
from duoformer_tcga_b200 import engine, ops  # noqa: E402
from duoformer_tcga_b200.graphs import GraphedForward  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1 << 21, 86 * 1160, 86 * 580, 86 * 290, 86 * 145, 86 * 95]
torch.manual_seed(0)
model = duo.build_model_no_extra_params(pretrained=False, depth=12, embed_dim=768, num_heads=12, num_classes=10,
                                        num_layers=4, proj_dim=768).cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
res = {}
with torch.no_grad():
    for _ in range(2):
        y_ref = model(x)
    for tokens in sizes:
        engine.SCALE_CHUNK_TOKENS = tokens
        ops.launch_count_reset()
        y = model(x)
        launches = ops.launch_count()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        eager_ms = (time.perf_counter() - t0) / 2 * 1e3
        g = GraphedForward(model, x, warmup=1)
        g(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            yg = g(x)
        e1.record()
        torch.cuda.synchronize()
        res[tokens] = {"chunk_tokens": tokens, "launches": launches, "eager_ms": round(eager_ms, 2),
                       "graph_ms": round(e0.elapsed_time(e1) / 3, 2),
                       "max_abs_diff_vs_default": float((yg.float() - y_ref.float()).abs().max())}
        print(json.dumps(res[tokens]), flush=True)
        del g
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_chunk.json", "w"), indent=1)
