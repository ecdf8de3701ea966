"""Does an HBM-bound kernel (LayerNorm, scale attention) hide under a tensor-bound GEMM on a
power-capped B200?  Runs each loop alone and then both on two streams; reports device time and
board power (NVML).  DUO_GEMM_MAX_SMS=n caps the GEMM grid."""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duoformer_tcga_b200 import ops  # noqa: E402


class Power:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.n = pynvml
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.s = []
        self.clk = []
        self.stop = False

    def __enter__(self):
        self.s, self.clk, self.stop = [], [], False
        self.t = threading.Thread(target=self.run)
        self.t.start()
        return self

    def run(self):
        while not self.stop:
            self.s.append(self.n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            self.clk.append(self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM))
            time.sleep(0.02)

    def __exit__(self, *a):
        self.stop = True
        self.t.join()

    def summary(self):
        s = sorted(self.s[len(self.s) // 3:]) or [0]
        c = sorted(self.clk[len(self.clk) // 3:]) or [0]
        return {"power_w_median": s[len(s) // 2], "sm_mhz_median": c[len(c) // 2]}


def main():
    images, D, S, P = 64, 768, 86, 49
    M = images * P * S
    dev = "cuda"
    A = (torch.randn(M, D, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(4 * D, D, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.zeros(4 * D, device=dev)
    hid = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
    x = torch.randn(M, D, device=dev)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    hn = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    qkv = torch.randn(M, 3 * D, device=dev).to(torch.bfloat16)
    ao = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    n_gemm = 300

    def gemm():
        ops.gemm(A, W, bias, hid, ops.EPI_GELU_BF16)

    def ln():
        ops.layernorm(x, g, b, hn, 1e-6)

    def attn():
        ops.group_attention(qkv, ao, S, 12, 0.125, algo=2)

    def run(name, f1, n1, f2, n2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with Power() as pw:
            e0.record()
            s1.wait_event(e0)
            s2.wait_event(e0)
            for i in range(max(n1, n2)):
                if f1 is not None and i < n1:
                    with torch.cuda.stream(s1):
                        f1()
                if f2 is not None and i < n2:
                    with torch.cuda.stream(s2):
                        f2()
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
        r = {"case": name, "ms": round(e0.elapsed_time(e1), 2), **pw.summary()}
        print(json.dumps(r), flush=True)
        return r

    for _ in range(3):
        gemm(), ln(), attn()
    res = []
    res.append(run("gemm_only", gemm, n_gemm, None, 0))
    res.append(run("ln_only", None, 0, ln, n_gemm))
    res.append(run("attn_only", None, 0, attn, n_gemm))
    res.append(run("gemm+ln", gemm, n_gemm, ln, n_gemm))
    res.append(run("gemm+attn", gemm, n_gemm, attn, n_gemm))
    res.append(run("gemm_only_again", gemm, n_gemm, None, 0))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/overlap_probe_{os.environ.get('DUO_GEMM_MAX_SMS', 'all')}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
