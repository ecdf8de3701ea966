"""Launch ONE GEMM shape of the scale block a few times (target for `ncu --set full --launch-skip 3 --launch-count 1`).
usage: one_gemm.py {qkv|qkv_ln|fc1_gelu|fc1_gelu_ln|proj|proj_fwd|fc2|fc2_fwd} [images]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duoformer_tcga_b200 import engine, ops  # noqa: E402

which = sys.argv[1]
images = int(sys.argv[2]) if len(sys.argv) > 2 else 64
D, M, dev = 768, images * 49 * 86, "cuda"
torch.manual_seed(0)
g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
x = torch.randn(M, D, device=dev) * 3
xb = x.to(torch.bfloat16)
parts = x.view(M, D // 256, 256)
pm = parts.mean(dim=2)
st = torch.stack([pm, ((parts - pm[:, :, None]) ** 2).sum(dim=2)], dim=2).contiguous()
N = {"qkv": 3 * D, "qkv_ln": 3 * D, "fc1_gelu": 4 * D, "fc1_gelu_ln": 4 * D}.get(which, D)
K = 4 * D if which.startswith("fc2") else D
W = torch.randn(N, K, device=dev) * 0.02
bias = torch.zeros(N, device=dev)
A = xb if K == D else (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
Wb = W.to(torch.bfloat16)
if which in ("qkv_ln", "fc1_gelu_ln"):
    w, bb = engine.pack_ln_linear(W, bias, g, b)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm(A, w, bb, out, ops.EPI_GELU_BF16 if "gelu" in which else ops.EPI_BF16, ln_stats=st)
elif which in ("qkv", "fc1_gelu"):
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm(A, Wb, bias, out, ops.EPI_GELU_BF16 if "gelu" in which else ops.EPI_BF16)
elif which.endswith("_fwd"):
    hn = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm(A, Wb, bias, x, ops.EPI_RESIDUAL_F32, xb_out=hn, stats_out=st)
else:
    fn = lambda: ops.gemm(A, Wb, bias, x, ops.EPI_RESIDUAL_F32)
for _ in range(6):
    fn()
torch.cuda.synchronize()
print("done", which, M, N, K)
