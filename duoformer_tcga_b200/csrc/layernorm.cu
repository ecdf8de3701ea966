// Fused LayerNorm (fp32 statistics, two-pass in registers), fp32 residual stream in,
// bf16 (or split hi|lo bf16) GEMM operand out.  HBM-bound: one warp per row, float4 loads,
// the row never leaves registers.  Replaces nn.LayerNorm(eps=1e-6) of
// scale_attention.py:65,78,91-92 / multiscale_attn.py:282-285.
#include "common.cuh"

namespace duo {
namespace {

constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// NV = dim / 128 float4 vectors per lane.
template <int NV, int OUT_KIND>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, void* __restrict__ out, int64_t rows, int64_t ldx,
                 float eps, float2* __restrict__ stats_out) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= rows) return;

  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = ld_stream_f4(xr + lane + 32 * i);

  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);

  // optional: (mean, sum of squared deviations) of every 256-column part, the layout duo_gemm's statistics forwarding
  // exchanges (a 256-column part = float4 vectors 64p .. 64p+63 = v[2p], v[2p+1] of every lane)
  if constexpr (NV % 2 == 0) {
    if (stats_out != nullptr) {
#pragma unroll
      for (int pt = 0; pt < NV / 2; ++pt) {
        const float4 a = v[2 * pt], b = v[2 * pt + 1];
        const float pm = warp_sum((a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w)) * (1.0f / 256.0f);
        float dq = (a.x - pm) * (a.x - pm) + (a.y - pm) * (a.y - pm) + (a.z - pm) * (a.z - pm) + (a.w - pm) * (a.w - pm);
        dq += (b.x - pm) * (b.x - pm) + (b.y - pm) * (b.y - pm) + (b.z - pm) * (b.z - pm) + (b.w - pm) * (b.w - pm);
        dq = warp_sum(dq);
        if (lane == 0) stats_out[row * (NV / 2) + pt] = make_float2(pm, dq);
      }
    }
  }

  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;  // float4 index inside the row
    const float4 g = __ldg(g4 + c4);
    const float4 b = __ldg(b4 + c4);
    const float y0 = (v[i].x - mean) * rstd * g.x + b.x;
    const float y1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float y2 = (v[i].z - mean) * rstd * g.z + b.z;
    const float y3 = (v[i].w - mean) * rstd * g.w + b.w;
    if constexpr (OUT_KIND == DUO_ACT_BF16) {
      uint2 w;
      w.x = pack_bf16x2(y0, y1);
      w.y = pack_bf16x2(y2, y3);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * D)[c4] = w;
    } else {
      uint2 h, l;
      pack_split2(y0, y1, h.x, l.x);
      pack_split2(y2, y3, h.y, l.y);
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row * (2 * D);
      reinterpret_cast<uint2*>(o)[c4] = h;
      reinterpret_cast<uint2*>(o + D)[c4] = l;
    }
  }
}

template <int NV>
int launch_ln(const float* x, const float* g, const float* b, void* out, int out_kind,
              int64_t rows, int64_t ldx, float eps, float2* stats, cudaStream_t st) {
  const int64_t grid = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  if (out_kind == DUO_ACT_BF16)
    layernorm_kernel<NV, DUO_ACT_BF16><<<static_cast<unsigned>(grid), kWarpsPerCta * 32, 0, st>>>(
        x, g, b, out, rows, ldx, eps, stats);
  else
    layernorm_kernel<NV, DUO_ACT_SPLIT><<<static_cast<unsigned>(grid), kWarpsPerCta * 32, 0, st>>>(
        x, g, b, out, rows, ldx, eps, stats);
  DUO_LAUNCH_CHECK("layernorm_kernel");
  return DUO_OK;
}

}  // namespace
}  // namespace duo

extern "C" int duo_layernorm(const float* x, const float* gamma, const float* beta, void* out,
                             int32_t out_kind, int64_t rows, int32_t dim, int64_t ldx,
                             float eps, float* stats_out, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(x && gamma && beta && out, "duo_layernorm: NULL pointer");
  DUO_CHECK_ARG(rows > 0, "duo_layernorm: rows=%lld", (long long)rows);
  DUO_CHECK_ARG(dim % 128 == 0 && dim >= 128 && dim <= 1024,
                "duo_layernorm: dim=%d must be a multiple of 128 in [128,1024]", dim);
  DUO_CHECK_ARG(out_kind == DUO_ACT_BF16 || out_kind == DUO_ACT_SPLIT,
                "duo_layernorm: out_kind=%d", out_kind);
  DUO_CHECK_ARG(ldx >= dim && ldx % 4 == 0, "duo_layernorm: ldx=%lld must be >= dim and a multiple of 4",
                (long long)ldx);
  DUO_CHECK_ARG((rows + kWarpsPerCta - 1) / kWarpsPerCta < (int64_t(1) << 31),
                "duo_layernorm: too many rows");
  DUO_CHECK_ARG(stats_out == nullptr || (dim % 256 == 0 && (reinterpret_cast<uintptr_t>(stats_out) & 7) == 0),
                "duo_layernorm: stats_out needs dim %% 256 == 0 and 8-byte alignment (dim=%d)", dim);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float2* stats = reinterpret_cast<float2*>(stats_out);
  switch (dim / 128) {
    case 1: return launch_ln<1>(x, gamma, beta, out, out_kind, rows, ldx, eps, stats, st);
    case 2: return launch_ln<2>(x, gamma, beta, out, out_kind, rows, ldx, eps, stats, st);
    case 3: return launch_ln<3>(x, gamma, beta, out, out_kind, rows, ldx, eps, stats, st);
    case 4: return launch_ln<4>(x, gamma, beta, out, out_kind, rows, ldx, eps, stats, st);
    case 5: return launch_ln<5>(x, gamma, beta, out, out_kind, rows, ldx, eps, stats, st);
    case 6: return launch_ln<6>(x, gamma, beta, out, out_kind, rows, ldx, eps, stats, st);
    case 7: return launch_ln<7>(x, gamma, beta, out, out_kind, rows, ldx, eps, stats, st);
    default: return launch_ln<8>(x, gamma, beta, out, out_kind, rows, ldx, eps, stats, st);
  }
}
