// Shared host/device helpers for libduoformer_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/duoformer_sm100.h"

namespace duo {

// ---- error plumbing (thread-local message, int codes across the ABI) -------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();

#define DUO_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      duo::set_error(__VA_ARGS__);          \
      return DUO_ERR_INVALID;               \
    }                                       \
  } while (0)

#define DUO_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return duo::cuda_fail(e__, #call); \
  } while (0)

// Checks the launch (not the execution: calls are asynchronous).
#define DUO_LAUNCH_CHECK(name)                                        \
  do {                                                                \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) return duo::cuda_fail(e__, name);         \
    duo::count_launch();                                              \
  } while (0)

int device_sm_count();
// true the first time it is called with `mask` on the current device (per-device one-time setup such as
// cudaFuncSetAttribute: function attributes are per device); false afterwards.
bool first_use_on_device(uint64_t& mask);

// ---- small device helpers --------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// GELU(erf) with erf evaluated by the fp32 rational minimax approximation x*P(x^2)/Q(x^2) on
// [-4, 4] (coefficients of the widely used Eigen/XLA float erf; max abs error 4.2e-7 vs erf,
// measured in numpy) — ~18 FMA-pipe ops + 1 MUFU.RCP per element instead of libdevice erff.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  float t = fminf(fmaxf(x * 0.70710678118654752440f, -4.0f), 4.0f);
  const float t2 = t * t;
  float p = -2.72614225801306e-10f;
  p = fmaf(p, t2, 2.77068142495902e-08f);
  p = fmaf(p, t2, -2.10102402082508e-06f);
  p = fmaf(p, t2, -5.69250639462346e-05f);
  p = fmaf(p, t2, -7.34990630326855e-04f);
  p = fmaf(p, t2, -2.95459980854025e-03f);
  p = fmaf(p, t2, -1.60960333262415e-02f);
  p *= t;
  float q = -1.45660718464996e-05f;
  q = fmaf(q, t2, -2.13374055278905e-04f);
  q = fmaf(q, t2, -1.68282697438203e-03f);
  q = fmaf(q, t2, -7.37332916720468e-03f);
  q = fmaf(q, t2, -1.42647390514189e-02f);
  const float e = __fdividef(p, q);
  const float hx = 0.5f * x;
  return fmaf(hx, e, hx);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2/FMUL2/FADD2: two fp32 lanes per issue slot) ----
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// gelu_erf_fast on two values at once: the Horner chains run as FFMA2.
__device__ __forceinline__ void gelu_erf_fast_x2(float& a, float& b) {
#define DUO_C2(v) pack2((v), (v))
  const uint64_t x = pack2(a, b);
  float ta, tb;
  unpack2(mul2(x, DUO_C2(0.70710678118654752440f)), ta, tb);
  ta = fminf(fmaxf(ta, -4.0f), 4.0f);
  tb = fminf(fmaxf(tb, -4.0f), 4.0f);
  const uint64_t t = pack2(ta, tb);
  const uint64_t t2 = mul2(t, t);
  uint64_t p = fma2(DUO_C2(-2.72614225801306e-10f), t2, DUO_C2(2.77068142495902e-08f));
  p = fma2(p, t2, DUO_C2(-2.10102402082508e-06f));
  p = fma2(p, t2, DUO_C2(-5.69250639462346e-05f));
  p = fma2(p, t2, DUO_C2(-7.34990630326855e-04f));
  p = fma2(p, t2, DUO_C2(-2.95459980854025e-03f));
  p = fma2(p, t2, DUO_C2(-1.60960333262415e-02f));
  p = mul2(p, t);
  uint64_t q = fma2(DUO_C2(-1.45660718464996e-05f), t2, DUO_C2(-2.13374055278905e-04f));
  q = fma2(q, t2, DUO_C2(-1.68282697438203e-03f));
  q = fma2(q, t2, DUO_C2(-7.37332916720468e-03f));
  q = fma2(q, t2, DUO_C2(-1.42647390514189e-02f));
  float qa, qb;
  unpack2(q, qa, qb);
  const uint64_t rq = pack2(rcp_approx(qa), rcp_approx(qb));
  const uint64_t e = mul2(p, rq);
  const uint64_t hx = mul2(x, DUO_C2(0.5f));
  unpack2(fma2(hx, e, hx), a, b);
#undef DUO_C2
}

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// GELU(erf) for the bf16 epilogue in "sigmoid of the probit logit" form:
//   gelu(x) = x * Phi(x) = x / (1 + exp(-L(x))),  L(x) = logit(Phi(x)) = ln(Phi(x) / Phi(-x)),
// L is odd and smooth; an odd degree-7 polynomial fitted (reweighted least squares, minimax on the
// gelu error, |x| clamped to 5.5 inside L only) reproduces the exact erf-GELU to 1.2e-5 absolute
// (fp32 simulation, x in [-12, 12]) — about 300x below bf16 rounding of O(1) activations — with
// 7 FMA-pipe operations + 2 MUFU (ex2, rcp) per element instead of ~18 + 1 for the rational erf.
// The coefficients below are -log2(e) * c_k so that exp(-L) = ex2(x * r(x^2)).
// Only u = x^2 is clamped (one FMNMX per element): beyond |x| = 5.5 the exponent continues linearly,
// x * r(30.25), so Phi still saturates to 0 / 1 (ex2 overflow -> rcp(inf) = 0 -> gelu = -0).
__device__ __forceinline__ uint64_t gelu_erf_sigmoid_p2(uint64_t x) {
#define DUO_C2(v) pack2((v), (v))
  float ua, ub;
  unpack2(mul2(x, x), ua, ub);
  const uint64_t u = pack2(fminf(ua, 30.25f), fminf(ub, 30.25f));
  uint64_t r = fma2(DUO_C2(2.4836352167767473e-05f), u, DUO_C2(7.360616000369191e-04f));
  r = fma2(r, u, DUO_C2(-1.0598272830247879e-01f));
  r = fma2(r, u, DUO_C2(-2.301647186279297f));
  float ea, eb;
  unpack2(mul2(x, r), ea, eb);
  const uint64_t d = add2(pack2(ex2_approx(ea), ex2_approx(eb)), DUO_C2(1.0f));
  float da, db;
  unpack2(d, da, db);
  return mul2(x, pack2(rcp_approx(da), rcp_approx(db)));
#undef DUO_C2
}
// Same function in tanh form (the GEMM epilogue's default): gelu(x) = hx + hx * tanh(x * q(x^2)), hx = x / 2, q an even
// degree-4 polynomial fitted (minimax on the gelu error) to 2.5e-5 absolute against the exact erf-GELU; ONE MUFU
// (tanh.approx, 2^-11 relative) per element instead of ex2 + rcp and one FMA less: 10 instead of 13 thread-instructions
// per pair.  Alone the two forms run at the same speed (0.931 vs 0.933 ms per 64 images); for the whole forward two ABAB
// sessions gave -1.5 % and 0.0 % (profiles/r02_notes.md).  q turns negative beyond |x| = 11.1, so x^2 is clamped at 100 (|x| = 10: tanh has saturated
// to +-1 in fp32 long before).
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint64_t gelu_erf_tanh_p2(uint64_t x) {
#define DUO_C2(v) pack2((v), (v))
  float ua, ub;
  unpack2(mul2(x, x), ua, ub);
  const uint64_t u = pack2(fminf(ua, 100.0f), fminf(ub, 100.0f));
  uint64_t r = fma2(DUO_C2(-3.51516791e-04f), u, DUO_C2(3.70056460e-02f));
  r = fma2(r, u, DUO_C2(7.97507884e-01f));
  float ea, eb;
  unpack2(mul2(x, r), ea, eb);
  const uint64_t t = pack2(tanh_approx(ea), tanh_approx(eb));
  const uint64_t hx = mul2(x, DUO_C2(0.5f));
  return fma2(hx, t, hx);
#undef DUO_C2
}
__device__ __forceinline__ void gelu_erf_sigmoid_x2(float& a, float& b) {
  unpack2(gelu_erf_sigmoid_p2(pack2(a, b)), a, b);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);  // .x = a (low 16 bits), .y = b
  return *reinterpret_cast<uint32_t*>(&v);
}

// hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ void pack_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat16 ah, al, bh, bl;
  split_bf16(a, ah, al);
  split_bf16(b, bh, bl);
  __nv_bfloat162 h = __halves2bfloat162(ah, bh), l = __halves2bfloat162(al, bl);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

}  // namespace duo
