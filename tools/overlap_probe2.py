"""Feasibility probe for SM-partitioned overlap: the four scale-block GEMMs looped on one stream (grid capped by
DUO_GEMM_MAX_SMS) and the block's HBM-bound kernels (LayerNorm, attention, LayerNorm) looped on a second stream,
in the real per-block ratio (64 images).  Reports each loop alone and both together."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duoformer_tcga_b200 import ops  # noqa: E402


def main():
    images, D, S, P = 64, 768, 86, 49
    M = images * P * S
    dev = "cuda"
    n = int(os.environ.get("PROBE_BLOCKS", "60"))
    bf = torch.bfloat16
    A = (torch.randn(M, D, device=dev) * 0.5).to(bf)
    Wqkv = (torch.randn(3 * D, D, device=dev) * 0.02).to(bf)
    Wproj = (torch.randn(D, D, device=dev) * 0.02).to(bf)
    W1 = (torch.randn(4 * D, D, device=dev) * 0.02).to(bf)
    W2 = (torch.randn(D, 4 * D, device=dev) * 0.02).to(bf)
    b3, b1, b4 = torch.zeros(3 * D, device=dev), torch.zeros(D, device=dev), torch.zeros(4 * D, device=dev)
    qkv_o = torch.empty(M, 3 * D, dtype=bf, device=dev)
    hid = torch.empty(M, 4 * D, dtype=bf, device=dev)
    x = torch.randn(M, D, device=dev)
    # second lane's own buffers
    x2 = torch.randn(M, D, device=dev)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    hn2 = torch.empty(M, D, dtype=bf, device=dev)
    qkv2 = torch.randn(M, 3 * D, device=dev).to(bf)
    ao2 = torch.empty(M, D, dtype=bf, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def gemms():
        ops.gemm(A, Wqkv, b3, qkv_o, ops.EPI_BF16)
        ops.gemm(A, Wproj, b1, x, ops.EPI_RESIDUAL_F32)
        ops.gemm(A, W1, b4, hid, ops.EPI_GELU_BF16)
        ops.gemm(hid, W2, b1, x, ops.EPI_RESIDUAL_F32)

    def hbm_ops():
        ops.layernorm(x2, g, b, hn2, 1e-6)
        ops.group_attention(qkv2, ao2, S, 12, 0.125)
        ops.layernorm(x2, g, b, hn2, 1e-6)

    def run(name, f1, f2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(n):
            if f1 is not None:
                with torch.cuda.stream(s1):
                    f1()
            if f2 is not None:
                with torch.cuda.stream(s2):
                    f2()
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        r = {"case": name, "ms_per_block": round(e0.elapsed_time(e1) / n, 4)}
        print(json.dumps(r), flush=True)
        return r

    for _ in range(3):
        gemms(), hbm_ops()
    res = [run("gemms_only", gemms, None), run("hbm_only", None, hbm_ops), run("both", gemms, hbm_ops),
           run("gemms_only_again", gemms, None), run("both_again", gemms, hbm_ops)]
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/overlap_probe2_{os.environ.get('DUO_GEMM_MAX_SMS', 'all')}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
