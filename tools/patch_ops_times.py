import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from duoformer_tcga_b200 import ops
B, N, D, H = 256, 50, 768, 12
rows = B * N
dev = "cuda"
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
Zs = torch.randn(rows, 2 * D, device=dev).to(torch.bfloat16)
Wq = torch.randn(3 * D, 2 * D, device=dev).to(torch.bfloat16) * 0.02
Wp = torch.randn(D, 2 * D, device=dev).to(torch.bfloat16) * 0.02
bq = torch.zeros(3 * D, device=dev); bp = torch.zeros(D, device=dev)
QKV32 = torch.empty(rows, 3 * D, device=dev)
AOs = torch.empty(rows, 2 * D, dtype=torch.bfloat16, device=dev)
Zo = torch.empty(rows, 2 * D, dtype=torch.bfloat16, device=dev)
print("qkv split3 -> f32      ", t(lambda: ops.gemm(Zs, Wq, bq, QKV32, ops.EPI_F32, split3=True)))
print("attn fp32 in, split out", t(lambda: ops.group_attention(QKV32, AOs, N, H, 0.125)))
QKVs = torch.empty(rows, 6 * D, dtype=torch.bfloat16, device=dev)
print("qkv split3 -> split    ", t(lambda: ops.gemm(Zs, Wq, bq, QKVs, ops.EPI_SPLIT_BF16, split3=True)))
print("attn split tcgen05     ", t(lambda: ops.group_attention(QKVs, AOs, N, H, 0.125, split_in=True)))
print("proj split3 -> split   ", t(lambda: ops.gemm(AOs, Wp, bp, Zo, ops.EPI_SPLIT_BF16, split3=True)))
Zb = Zs[:, :D].contiguous(); Wqb = Wq[:, :D].contiguous(); Wpb = Wp[:, :D].contiguous()
QKVb = torch.empty(rows, 3 * D, dtype=torch.bfloat16, device=dev); AOb = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
print("qkv bf16               ", t(lambda: ops.gemm(Zb, Wqb, bq, QKVb, ops.EPI_BF16)))
print("attn bf16 mma          ", t(lambda: ops.group_attention(QKVb, AOb, N, H, 0.125)))
print("proj bf16              ", t(lambda: ops.gemm(AOb, Wpb, bp, Zb, ops.EPI_BF16)))
