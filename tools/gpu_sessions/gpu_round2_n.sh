#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/n_summary.txt; : > $S
DUO_TEST_LIB=duoformer_tcga_b200/libduoformer_sm100_fwd_gelu1buf.so timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "gemm" > gpurun_out/n_kernels_1buf.log 2>&1; echo "kernels_1buf rc=$?" >> $S
DUO_TEST_LIB=duoformer_tcga_b200/libduoformer_sm100_fwd_stef.so timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "gemm" > gpurun_out/n_kernels_stef.log 2>&1; echo "kernels_stef rc=$?" >> $S
for rep in 1 2; do for v in "" _fwd_gelu1buf _fwd_stef; do
  timeout 300 python tools/bench_kernels.py --images 64 --only fc1,fwd,qkv --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag _n$rep$v > gpurun_out/n${rep}_bench_kernels$v.log 2>&1; echo "bench_kernels$rep$v rc=$?" >> $S
done; done
cat $S; tail -2 gpurun_out/n_kernels_1buf.log; tail -2 gpurun_out/n_kernels_stef.log
for rep in 1 2; do for v in "" _fwd_gelu1buf _fwd_stef; do echo "== $rep lib$v"; grep -E "gemm_fc1_gelu|gemm_qkv|ln_applied|seq_.*forwarding" gpurun_out/n${rep}_bench_kernels$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['kernel'].ljust(34), d['ms'], d.get('tflops'))"; done; done
