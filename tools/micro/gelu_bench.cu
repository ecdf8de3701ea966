// Microbenchmark: throughput of GELU(erf) formulations on sm_100a (elements / clk / SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../duoformer_tcga_b200/csrc -o gelu_bench gelu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace duo;

__device__ __forceinline__ float gelu_poly(float x) {  // placeholder cheap variant: deg-5 odd poly (NOT accurate)
  float t = fminf(fmaxf(x, -4.f), 4.f), t2 = t * t;
  float r = 1e-4f; r = fmaf(r, t2, -2e-3f); r = fmaf(r, t2, 2e-2f); r = fmaf(r, t2, -6e-2f); r = fmaf(r, t2, 0.39f);
  return x * fmaf(t, r, 0.5f);
}

template <int V>
__global__ void k(const float* in, float* out, int iters) {
  float x[32];
  for (int j = 0; j < 32; ++j) x[j] = in[threadIdx.x + 32 * j];
  for (int it = 0; it < iters; ++it) {
    if (V == 0) { for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]) + 0.5f; }
    if (V == 1) { for (int j = 0; j < 32; ++j) x[j] = gelu_erf_fast(x[j]) + 0.5f; }
    if (V == 2) { for (int j = 0; j < 32; j += 2) { gelu_erf_fast_x2(x[j], x[j + 1]); x[j] += 0.5f; x[j+1] += 0.5f; } }
    if (V == 4) { for (int j = 0; j < 32; j += 2) { gelu_erf_sigmoid_x2(x[j], x[j + 1]); x[j] += 0.5f; x[j+1] += 0.5f; } }
    if (V == 3) { for (int j = 0; j < 32; ++j) x[j] = gelu_poly(x[j]) + 0.5f; }
  }
  float s = 0; for (int j = 0; j < 32; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V> void run(const char* name, float* in, float* out, int warps_per_sm) {
  int iters = 2000; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<V><<<148, 32 * warps_per_sm>>>(in, out, 10); cudaDeviceSynchronize();
  cudaEventRecord(a); k<V><<<148, 32 * warps_per_sm>>>(in, out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double elems = 148.0 * warps_per_sm * 32 * 32 * iters;
  printf("%-22s warps/SM=%2d  %.3f ms  %.1f Gelem/s  (%.2f elem/ns/SM)\n", name, warps_per_sm, ms, elems / ms / 1e6, elems / ms / 1e6 / 148);
}
int main() {
  float *in, *out; cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMemset(in, 0, 4096 * 4);
  for (int w : {4, 8, 16}) {
    run<0>("erff", in, out, w); run<1>("rational scalar", in, out, w); run<2>("rational packed x2", in, out, w); run<3>("poly5 scalar", in, out, w); run<4>("sigmoid packed x2", in, out, w);
  }
  return 0;
}
