import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from duoformer_tcga_b200 import ops
S, H, G = 86, 12, 256 * 49
qkv = torch.randn(G * S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(G, H * 64, dtype=torch.bfloat16, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("q_rows=1 attention, B=256:", t(lambda: ops.group_attention(qkv, out, S, H, 0.125, q_rows=1)))
