"""Per-kernel microbenchmarks on one B200 (CUDA events, L2-flushing working sets).

Prints one JSON line per kernel with achieved TFLOP/s or GB/s and the fraction of the
measured peaks in MEASURED_PEAKS.json (fallback: 1590 TF / 6650 GB/s)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duoformer_tcga_b200 import ops  # noqa: E402


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d["hbm_gbs"], "measured"
    return 1590.0, 6650.0, "fallback"


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--S", type=int, default=86)
    ap.add_argument("--only", default="")
    ap.add_argument("--lib", default="", help="alternative build of the library (csrc/Makefile `tuning` targets)")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    if a.lib:
        from duoformer_tcga_b200 import _lib
        _lib.LIB_PATH = os.path.abspath(a.lib)
    tf_peak, gb_peak, which = peaks()
    D, P, S = 768, 49, a.S
    M = a.images * P * S
    dev = "cuda"
    res = []

    def emit(name, ms, best, flops=None, nbytes=None):
        r = {"kernel": name, "M": M, "ms": round(ms, 4), "best_ms": round(best, 4)}
        if flops:
            r["tflops"] = round(flops / ms / 1e9, 1)
            r["frac_tensor_peak"] = round(flops / ms / 1e9 / tf_peak, 3)
        if nbytes:
            r["gbs"] = round(nbytes / ms / 1e6, 1)
            r["frac_hbm_peak"] = round(nbytes / ms / 1e6 / gb_peak, 3)
        r["peaks"] = which
        print(json.dumps(r), flush=True)
        res.append(r)

    x = torch.randn(M, D, device=dev) * 3
    g = torch.ones(D, device=dev)
    b = torch.zeros(D, device=dev)
    hn = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    if not a.only or "ln" in a.only:
        ms, best = timeit(lambda: ops.layernorm(x, g, b, hn, 1e-6))
        emit("layernorm_f32_to_bf16", ms, best, nbytes=M * D * 6)

    shapes = [("qkv", 3 * D, D, ops.EPI_BF16), ("proj_residual", D, D, ops.EPI_RESIDUAL_F32),
              ("fc1_gelu", 4 * D, D, ops.EPI_GELU_BF16), ("fc1_nogelu", 4 * D, D, ops.EPI_BF16),
              ("fc2_residual", D, 4 * D, ops.EPI_RESIDUAL_F32), ("fc2_shape_bf16out", D, 4 * D, ops.EPI_BF16),
              ("proj_shape_bf16out", D, D, ops.EPI_BF16)]
    for name, N, K, epi in shapes:
        if a.only and "gemm" not in a.only and not any(tok in name for tok in a.only.split(",")):
            continue
        A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        W = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        bias = torch.zeros(N, device=dev)
        out = x if epi == ops.EPI_RESIDUAL_F32 else torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ms, best = timeit(lambda: ops.gemm(A, W, bias, out, epi))
        emit("gemm_" + name, ms, best, flops=2.0 * M * N * K)
        # library bar: torch.matmul (cuBLAS) on the same operands, no epilogue
        ms2, best2 = timeit(lambda: torch.matmul(A, W.t()))
        emit("cublas_" + name, ms2, best2, flops=2.0 * M * N * K)
        del A, W, out

    # LayerNorm statistics forwarding: residual GEMM (+ bf16 copy + statistics) -> consumer GEMM (LayerNorm in the
    # epilogue) against the three-launch sequence residual GEMM, LayerNorm, GEMM
    if not a.only or "fwd" in a.only:
        from duoformer_tcga_b200 import engine
        st = torch.empty(M, D // 256, 2, device=dev)
        for name, K, N2, epi2 in (("proj", D, 4 * D, ops.EPI_GELU_BF16), ("fc2", 4 * D, 3 * D, ops.EPI_BF16)):
            A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
            W = (torch.randn(D, K, device=dev) * 0.02).to(torch.bfloat16)
            bias = torch.zeros(D, device=dev)
            W2 = torch.randn(N2, D, device=dev) * 0.02
            b2 = torch.zeros(N2, device=dev)
            w2p, b2p = engine.pack_ln_linear(W2, b2, g, b)
            W2b = W2.to(torch.bfloat16)
            out2 = torch.empty(M, N2, dtype=torch.bfloat16, device=dev)
            ms, best = timeit(lambda: ops.gemm(A, W, bias, x, ops.EPI_RESIDUAL_F32, xb_out=hn, stats_out=st))
            emit(f"gemm_{name}_residual_fwd", ms, best, flops=2.0 * M * D * K, nbytes=M * (K * 2 + D * 10))
            ms, best = timeit(lambda: ops.gemm(A, W, bias, x, ops.EPI_RESIDUAL_F32))
            emit(f"gemm_{name}_residual", ms, best, flops=2.0 * M * D * K, nbytes=M * (K * 2 + D * 8))
            ms, best = timeit(lambda: ops.gemm(hn, w2p, b2p, out2, epi2, ln_stats=st))
            emit(f"gemm_after_{name}_ln_applied", ms, best, flops=2.0 * M * N2 * D)
            ms, best = timeit(lambda: ops.gemm(hn, W2b, b2, out2, epi2))
            emit(f"gemm_after_{name}_plain", ms, best, flops=2.0 * M * N2 * D)

            def fwd_seq():
                ops.gemm(A, W, bias, x, ops.EPI_RESIDUAL_F32, xb_out=hn, stats_out=st)
                ops.gemm(hn, w2p, b2p, out2, epi2, ln_stats=st)

            def ln_seq():
                ops.gemm(A, W, bias, x, ops.EPI_RESIDUAL_F32)
                ops.layernorm(x, g, b, hn, 1e-6)
                ops.gemm(hn, W2b, b2, out2, epi2)
            ms, best = timeit(fwd_seq)
            emit(f"seq_{name}_forwarding", ms, best, flops=2.0 * M * D * (K + N2))
            ms, best = timeit(ln_seq)
            emit(f"seq_{name}_layernorm_launch", ms, best, flops=2.0 * M * D * (K + N2))
            del A, W, W2, W2b, w2p, out2

    # token scatter GEMMs (projection of one trunk stage straight into token rows + positional add)
    if not a.only or "scatter" in a.only or "gemm" in a.only:
        from duoformer_tcga_b200.index_tables import token_row_maps
        maps = token_row_maps(4, 7)
        pos = torch.randn(S, D, device=dev)
        for k, (C, hw) in {0: (256, 56), 1: (512, 28), 2: (1024, 14), 3: (2048, 7)}.items():
            if k not in maps or S != 86:
                continue
            rows = a.images * hw * hw
            A = (torch.randn(rows, C, device=dev) * 0.5).to(torch.bfloat16)
            W = (torch.randn(D, C, device=dev) * 0.02).to(torch.bfloat16)
            bias = torch.zeros(D, device=dev)
            rm = maps[k].to(dev)
            ms, best = timeit(lambda: ops.gemm(A, W, bias, x, ops.EPI_SCATTER_F32, row_map=rm, rows_per_group=hw * hw,
                                               dest_rows_per_group=P * S, pos=pos, pos_period=S))
            emit(f"gemm_scatter_K{C}", ms, best, flops=2.0 * rows * D * C, nbytes=rows * (C * 2 + D * 4))
            del A, W

    qkv = (torch.randn(M, 3 * D, device=dev)).to(torch.bfloat16)
    ao = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    for algo in (4, 3, 2, 1):
        if a.only and "attn" not in a.only:
            continue
        if algo == 4 and S > 8:
            continue
        if algo == 2 and not (16 < S <= 96):
            continue
        if algo == 3 and not (64 < S <= 96):
            continue
        ms, best = timeit(lambda: ops.group_attention(qkv, ao, S, 12, 0.125, algo=algo), iters=5, warmup=2)
        emit(f"scale_attention_algo{algo}_S{S}", ms, best, flops=4.0 * (M // S) * S * S * D, nbytes=M * D * 8)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    json.dump(res, open(os.path.join(out_dir, f"bench_kernels{a.tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
