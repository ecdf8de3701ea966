// Grouped multi-head attention over S consecutive rows of a fused qkv matrix (head_dim 64).
//
//   scale attention : group = one patch, S = 6 / 22 / 86 scale tokens
//                     (scale_attention.py:28-45, multiscale_attn.py:149-166)
//   patch attention : group = one image, S = P+1 tokens
//                     (scale_attention.py:195-207, multiscale_attn.py:205-216)
//
// Two kernels:
//   algo 1  warp-per-(group, head) register/shuffle kernel: K and V of the head live in shared
//           memory, each lane owns ceil(S/32) keys, softmax statistics via warp shuffles, all
//           arithmetic fp32 FMA.  Exact enough for the fp32 mode; HBM-bound for small S.
//   algo 2  warp-level tensor-core kernel (mma.sync m16n8k16 bf16, fp32 accumulate) for
//           16 < S <= 96: one CTA of two warps per (group, head), Q/K/V staged once in
//           XOR-swizzled shared memory via cp.async, scores/probabilities stay in registers
//           (the accumulator fragment of QK^T is re-used as the A fragment of PV).
#include "common.cuh"

namespace duo {
namespace {

constexpr int kHeadDim = 64;

// =============================== algo 1: FMA ===============================================
template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<__nv_bfloat16> {
  static constexpr int kWordsPerRow = 32;  // 64 bf16
};
template <>
struct ElemTraits<float> {
  static constexpr int kWordsPerRow = 64;
};

template <typename Tin>
__host__ __device__ constexpr int fma_smem_words_per_warp(int S, int kpl) {
  // V rows dense, q row, p row, then K rows padded by one word (conflict-free column reads);
  // rounded up so every warp's region stays 16-byte aligned.
  const int words =
      S * ElemTraits<Tin>::kWordsPerRow + 64 + kpl * 32 + S * (ElemTraits<Tin>::kWordsPerRow + 1);
  return (words + 3) & ~3;
}

// One warp per (group, head), several problems per CTA: small S, or a single query row.
// (S >= 32 with more than one query row goes to the cooperative, query-blocked kernel below.)
template <typename Tin, int OUT_KIND, int KPL>
__global__ void group_attention_fma_kernel(const Tin* __restrict__ qkv, void* __restrict__ out,
                                           int64_t num_problems, int S, int H, float scale,
                                           int q_rows) {
  constexpr int WPR = ElemTraits<Tin>::kWordsPerRow;
  constexpr int VPR = WPR / 4;          // 16-byte vectors per row
  constexpr int RPP = 32 / VPR;         // rows loaded per warp pass
  constexpr bool kIsBf16 = (WPR == 32);
  extern __shared__ uint32_t smem_words[];

  const int warps_per_cta = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t prob = static_cast<int64_t>(blockIdx.x) * warps_per_cta + warp;
  if (prob >= num_problems) return;
  const int64_t g = prob / H;
  const int h = static_cast<int>(prob - g * H);
  const int D = H * kHeadDim;
  const int64_t ld = 3 * static_cast<int64_t>(D);

  uint32_t* Vs = smem_words + static_cast<size_t>(warp) * fma_smem_words_per_warp<Tin>(S, KPL);
  float* qs = reinterpret_cast<float*>(Vs + S * WPR);
  float* ps = qs + 64;
  uint32_t* Ks = reinterpret_cast<uint32_t*>(ps + KPL * 32);

  const Tin* base = qkv + (g * S) * ld + h * kHeadDim;
  // ---- stage K and V of this head ----
  const int stage_tid = lane;
  const int stage_rows = 32 / VPR;
  for (int r0 = 0; r0 < S; r0 += stage_rows) {
    const int r = r0 + stage_tid / VPR;
    const int vec = stage_tid % VPR;
    if (r < S) {
      const uint4 kv = __ldg(reinterpret_cast<const uint4*>(base + r * ld + D) + vec);
      const uint4 vv = __ldg(reinterpret_cast<const uint4*>(base + r * ld + 2 * D) + vec);
      uint32_t* kd = Ks + r * (WPR + 1) + vec * 4;
      kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
      uint32_t* vd = Vs + r * WPR + vec * 4;
      vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
    }
  }
  __syncwarp();

  int krow[KPL];
#pragma unroll
  for (int kk = 0; kk < KPL; ++kk) {
    const int j = lane + 32 * kk;
    krow[kk] = (j < S ? j : S - 1) * (WPR + 1);
  }

  for (int i = 0; i < q_rows; ++i) {
    // q row -> shared (fp32)
    {
      const Tin* qp = base + i * ld + 2 * lane;
      float q0, q1;
      if constexpr (kIsBf16) {
        const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(qp);
        q0 = __bfloat162float(q2.x);
        q1 = __bfloat162float(q2.y);
      } else {
        const float2 q2 = *reinterpret_cast<const float2*>(qp);
        q0 = q2.x;
        q1 = q2.y;
      }
      qs[2 * lane] = q0;
      qs[2 * lane + 1] = q1;
    }
    __syncwarp();

    float acc[KPL];
#pragma unroll
    for (int kk = 0; kk < KPL; ++kk) acc[kk] = 0.f;
    if constexpr (kIsBf16) {
#pragma unroll 8
      for (int w = 0; w < 32; w += 2) {
        const float4 q4 = *reinterpret_cast<const float4*>(qs + 2 * w);
#pragma unroll
        for (int kk = 0; kk < KPL; ++kk) {
          const uint32_t k01 = Ks[krow[kk] + w];
          const uint32_t k23 = Ks[krow[kk] + w + 1];
          acc[kk] = fmaf(q4.x, __uint_as_float(k01 << 16), acc[kk]);
          acc[kk] = fmaf(q4.y, __uint_as_float(k01 & 0xffff0000u), acc[kk]);
          acc[kk] = fmaf(q4.z, __uint_as_float(k23 << 16), acc[kk]);
          acc[kk] = fmaf(q4.w, __uint_as_float(k23 & 0xffff0000u), acc[kk]);
        }
      }
    } else {
#pragma unroll 8
      for (int w = 0; w < 64; w += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(qs + w);
#pragma unroll
        for (int kk = 0; kk < KPL; ++kk) {
          const float* kr = reinterpret_cast<const float*>(Ks) + krow[kk] + w;
          acc[kk] = fmaf(q4.x, kr[0], acc[kk]);
          acc[kk] = fmaf(q4.y, kr[1], acc[kk]);
          acc[kk] = fmaf(q4.z, kr[2], acc[kk]);
          acc[kk] = fmaf(q4.w, kr[3], acc[kk]);
        }
      }
    }

    float m = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < KPL; ++kk) {
      acc[kk] = (lane + 32 * kk < S) ? acc[kk] * scale : -INFINITY;
      m = fmaxf(m, acc[kk]);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int kk = 0; kk < KPL; ++kk) {
      acc[kk] = (lane + 32 * kk < S) ? __expf(acc[kk] - m) : 0.f;
      sum += acc[kk];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int kk = 0; kk < KPL; ++kk) ps[lane + 32 * kk] = acc[kk] * inv;
    __syncwarp();

    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = ps[j];
      if constexpr (kIsBf16) {
        const uint32_t v01 = Vs[j * WPR + lane];
        o0 = fmaf(pj, __uint_as_float(v01 << 16), o0);
        o1 = fmaf(pj, __uint_as_float(v01 & 0xffff0000u), o1);
      } else {
        const float2 v01 = *reinterpret_cast<const float2*>(Vs + j * WPR + 2 * lane);
        o0 = fmaf(pj, v01.x, o0);
        o1 = fmaf(pj, v01.y, o1);
      }
    }

    const int64_t orow = g * q_rows + i;
    const int ocol = h * kHeadDim + 2 * lane;
    if constexpr (OUT_KIND == DUO_ACT_BF16) {
      reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * D + ocol)[0] =
          pack_bf16x2(o0, o1);
    } else if constexpr (OUT_KIND == DUO_ACT_SPLIT) {
      uint32_t hi, lo;
      pack_split2(o0, o1, hi, lo);
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + orow * (2 * D) + ocol;
      reinterpret_cast<uint32_t*>(o)[0] = hi;
      reinterpret_cast<uint32_t*>(o + D)[0] = lo;
    } else {
      reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + orow * D + ocol)[0] =
          make_float2(o0, o1);
    }
    __syncwarp();
  }
}

// ---- cooperative variant with query blocking (S >= 32, more than one query row) -------------
// The single-query loop above is bound by shared-memory wavefronts (every K / V word is re-read for
// every query: ~300 wavefronts per query at S = 50).  Here a warp takes kQB queries at a time, so a
// K row or V row read from shared memory feeds kQB dot products, and the probabilities are fetched
// four keys per broadcast load: ~85 wavefronts per query.
constexpr int kQB = 4;

template <typename Tin, int OUT_KIND, int KPL>
__global__ void group_attention_fma_qb_kernel(const Tin* __restrict__ qkv, void* __restrict__ out,
                                              int S, int H, float scale, int q_rows) {
  constexpr int WPR = ElemTraits<Tin>::kWordsPerRow;
  constexpr int VPR = WPR / 4;
  constexpr bool kIsBf16 = (WPR == 32);
  constexpr int kPS = KPL * 32;  // probability row length
  extern __shared__ uint32_t smem_words[];

  const int warps_per_cta = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t prob = blockIdx.x;
  const int64_t g = prob / H;
  const int h = static_cast<int>(prob - g * H);
  const int D = H * kHeadDim;
  const int64_t ld = 3 * static_cast<int64_t>(D);

  uint32_t* Vs = smem_words;
  uint32_t* Ks = Vs + S * WPR;
  float* scratch = reinterpret_cast<float*>(smem_words + ((S * WPR + S * (WPR + 1) + 3) & ~3));
  float* qs = scratch + warp * (kQB * 64 + kQB * kPS);
  float* ps = qs + kQB * 64;

  const Tin* base = qkv + (g * S) * ld + h * kHeadDim;
  {
    const int stage_rows = static_cast<int>(blockDim.x) / VPR;
    for (int r0 = 0; r0 < S; r0 += stage_rows) {
      const int r = r0 + static_cast<int>(threadIdx.x) / VPR;
      const int vec = static_cast<int>(threadIdx.x) % VPR;
      if (r < S) {
        const uint4 kv = __ldg(reinterpret_cast<const uint4*>(base + r * ld + D) + vec);
        const uint4 vv = __ldg(reinterpret_cast<const uint4*>(base + r * ld + 2 * D) + vec);
        uint32_t* kd = Ks + r * (WPR + 1) + vec * 4;
        kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
        uint32_t* vd = Vs + r * WPR + vec * 4;
        vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
      }
    }
  }
  __syncthreads();

  int krow[KPL];
#pragma unroll
  for (int kk = 0; kk < KPL; ++kk) {
    const int j = lane + 32 * kk;
    krow[kk] = (j < S ? j : S - 1) * (WPR + 1);
  }

  for (int i0 = warp * kQB; i0 < q_rows; i0 += warps_per_cta * kQB) {
    // kQB query rows -> shared (fp32); rows past q_rows repeat the last one (results discarded)
#pragma unroll
    for (int qb = 0; qb < kQB; ++qb) {
      const int i = (i0 + qb < q_rows) ? i0 + qb : q_rows - 1;
      const Tin* qp = base + i * ld + 2 * lane;
      float q0, q1;
      if constexpr (kIsBf16) {
        const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(qp);
        q0 = __bfloat162float(q2.x);
        q1 = __bfloat162float(q2.y);
      } else {
        const float2 q2 = *reinterpret_cast<const float2*>(qp);
        q0 = q2.x;
        q1 = q2.y;
      }
      qs[qb * 64 + 2 * lane] = q0;
      qs[qb * 64 + 2 * lane + 1] = q1;
    }
    __syncwarp();

    float acc[KPL][kQB];
#pragma unroll
    for (int kk = 0; kk < KPL; ++kk)
#pragma unroll
      for (int qb = 0; qb < kQB; ++qb) acc[kk][qb] = 0.f;
#pragma unroll 4
    for (int e = 0; e < 64; e += 4) {  // four head-dim elements per step
      float4 q4[kQB];
#pragma unroll
      for (int qb = 0; qb < kQB; ++qb) q4[qb] = *reinterpret_cast<const float4*>(qs + qb * 64 + e);
#pragma unroll
      for (int kk = 0; kk < KPL; ++kk) {
        float k0, k1, k2, k3;
        if constexpr (kIsBf16) {
          const uint32_t k01 = Ks[krow[kk] + (e >> 1)];
          const uint32_t k23 = Ks[krow[kk] + (e >> 1) + 1];
          k0 = __uint_as_float(k01 << 16);
          k1 = __uint_as_float(k01 & 0xffff0000u);
          k2 = __uint_as_float(k23 << 16);
          k3 = __uint_as_float(k23 & 0xffff0000u);
        } else {
          const float* kr = reinterpret_cast<const float*>(Ks) + krow[kk] + e;
          k0 = kr[0]; k1 = kr[1]; k2 = kr[2]; k3 = kr[3];
        }
#pragma unroll
        for (int qb = 0; qb < kQB; ++qb) {
          acc[kk][qb] = fmaf(q4[qb].x, k0, acc[kk][qb]);
          acc[kk][qb] = fmaf(q4[qb].y, k1, acc[kk][qb]);
          acc[kk][qb] = fmaf(q4[qb].z, k2, acc[kk][qb]);
          acc[kk][qb] = fmaf(q4[qb].w, k3, acc[kk][qb]);
        }
      }
    }

#pragma unroll
    for (int qb = 0; qb < kQB; ++qb) {
      float m = -INFINITY;
#pragma unroll
      for (int kk = 0; kk < KPL; ++kk) {
        acc[kk][qb] = (lane + 32 * kk < S) ? acc[kk][qb] * scale : -INFINITY;
        m = fmaxf(m, acc[kk][qb]);
      }
      m = warp_max(m);
      float sum = 0.f;
#pragma unroll
      for (int kk = 0; kk < KPL; ++kk) {
        acc[kk][qb] = (lane + 32 * kk < S) ? __expf(acc[kk][qb] - m) : 0.f;
        sum += acc[kk][qb];
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int kk = 0; kk < KPL; ++kk) ps[qb * kPS + lane + 32 * kk] = acc[kk][qb] * inv;
    }
    __syncwarp();

    float o0[kQB], o1[kQB];
#pragma unroll
    for (int qb = 0; qb < kQB; ++qb) o0[qb] = o1[qb] = 0.f;
    auto v_pair = [&](int j, float& v0, float& v1) {
      if constexpr (kIsBf16) {
        const uint32_t v01 = Vs[j * WPR + lane];
        v0 = __uint_as_float(v01 << 16);
        v1 = __uint_as_float(v01 & 0xffff0000u);
      } else {
        const float2 v01 = *reinterpret_cast<const float2*>(Vs + j * WPR + 2 * lane);
        v0 = v01.x;
        v1 = v01.y;
      }
    };
    const int S4 = S & ~3;
    for (int j = 0; j < S4; j += 4) {
      float4 p4[kQB];
#pragma unroll
      for (int qb = 0; qb < kQB; ++qb) p4[qb] = *reinterpret_cast<const float4*>(ps + qb * kPS + j);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float v0, v1;
        v_pair(j + t, v0, v1);
#pragma unroll
        for (int qb = 0; qb < kQB; ++qb) {
          const float pj = t == 0 ? p4[qb].x : (t == 1 ? p4[qb].y : (t == 2 ? p4[qb].z : p4[qb].w));
          o0[qb] = fmaf(pj, v0, o0[qb]);
          o1[qb] = fmaf(pj, v1, o1[qb]);
        }
      }
    }
    for (int j = S4; j < S; ++j) {
      float v0, v1;
      v_pair(j, v0, v1);
#pragma unroll
      for (int qb = 0; qb < kQB; ++qb) {
        const float pj = ps[qb * kPS + j];
        o0[qb] = fmaf(pj, v0, o0[qb]);
        o1[qb] = fmaf(pj, v1, o1[qb]);
      }
    }

    const int ocol = h * kHeadDim + 2 * lane;
#pragma unroll
    for (int qb = 0; qb < kQB; ++qb) {
      if (i0 + qb >= q_rows) break;
      const int64_t orow = g * q_rows + i0 + qb;
      if constexpr (OUT_KIND == DUO_ACT_BF16) {
        reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * D + ocol)[0] =
            pack_bf16x2(o0[qb], o1[qb]);
      } else if constexpr (OUT_KIND == DUO_ACT_SPLIT) {
        uint32_t hi, lo;
        pack_split2(o0[qb], o1[qb], hi, lo);
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + orow * (2 * D) + ocol;
        reinterpret_cast<uint32_t*>(o)[0] = hi;
        reinterpret_cast<uint32_t*>(o + D)[0] = lo;
      } else {
        reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + orow * D + ocol)[0] =
            make_float2(o0[qb], o1[qb]);
      }
    }
    __syncwarp();
  }
}

template <typename Tin, int OUT_KIND, int KPL>
int launch_fma(const void* qkv, void* out, int64_t groups, int S, int H, float scale, int q_rows,
               cudaStream_t st) {
  const int64_t problems = groups * H;
  if (S >= 32 && q_rows > 1) {
    // cooperative CTA per problem: one K/V copy shared by four warps, kQB queries per warp pass
    constexpr int WPR = ElemTraits<Tin>::kWordsPerRow;
#ifndef DUO_FMA_QB_WARPS
#define DUO_FMA_QB_WARPS 8  // warps sharing one staged K / V copy (4: latency-bound at S = 145, 0.87 ms per block of config 4)
#endif
    const int warps = DUO_FMA_QB_WARPS;
    const size_t smem =
        (static_cast<size_t>((S * WPR + S * (WPR + 1) + 3) & ~3) + warps * (kQB * 64 + kQB * KPL * 32)) * 4;
    if (smem > 220 * 1024 || problems >= (int64_t(1) << 31)) {
      set_error("duo_group_attention: S=%d / %lld problems unsupported", S, (long long)problems);
      return DUO_ERR_INVALID;
    }
    auto kfn = group_attention_fma_qb_kernel<Tin, OUT_KIND, KPL>;
    DUO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kfn<<<static_cast<unsigned>(problems), warps * 32, smem, st>>>(reinterpret_cast<const Tin*>(qkv), out, S, H, scale,
                                                                   q_rows);
    DUO_LAUNCH_CHECK("group_attention_fma_qb_kernel");
    return DUO_OK;
  }
  const size_t per_warp = static_cast<size_t>(fma_smem_words_per_warp<Tin>(S, KPL)) * 4;
  int warps = 4;
  while (warps > 1 && per_warp * warps > 100 * 1024) warps >>= 1;
  const size_t smem = per_warp * warps;
  if (smem > 220 * 1024) {
    set_error("duo_group_attention: S=%d needs %zu B of shared memory", S, smem);
    return DUO_ERR_INVALID;
  }
  auto kfn = group_attention_fma_kernel<Tin, OUT_KIND, KPL>;
  DUO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const int64_t grid = (problems + warps - 1) / warps;
  if (grid >= (int64_t(1) << 31)) {
    set_error("duo_group_attention: too many groups");
    return DUO_ERR_INVALID;
  }
  kfn<<<static_cast<unsigned>(grid), warps * 32, smem, st>>>(reinterpret_cast<const Tin*>(qkv), out,
                                                            problems, S, H, scale, q_rows);
  DUO_LAUNCH_CHECK("group_attention_fma_kernel");
  return DUO_OK;
}

template <typename Tin, int OUT_KIND>
int dispatch_fma_kpl(const void* qkv, void* out, int64_t groups, int S, int H, float scale, int q_rows,
                     cudaStream_t st) {
  const int kpl = (S + 31) / 32;
  switch (kpl) {
    case 1: return launch_fma<Tin, OUT_KIND, 1>(qkv, out, groups, S, H, scale, q_rows, st);
    case 2: return launch_fma<Tin, OUT_KIND, 2>(qkv, out, groups, S, H, scale, q_rows, st);
    case 3: return launch_fma<Tin, OUT_KIND, 3>(qkv, out, groups, S, H, scale, q_rows, st);
    case 4: return launch_fma<Tin, OUT_KIND, 4>(qkv, out, groups, S, H, scale, q_rows, st);
    case 5: return launch_fma<Tin, OUT_KIND, 5>(qkv, out, groups, S, H, scale, q_rows, st);
    default: set_error("duo_group_attention: S=%d > 160 unsupported", S); return DUO_ERR_INVALID;
  }
}

// =============================== algo 2: mma.sync ==========================================
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                            uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                                  uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Byte offset of 16-byte chunk `c` (0..7) of row `r` in a [rows][128 B] XOR-swizzled tile.
__device__ __forceinline__ uint32_t swz(int r, int c) {
  return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4));
}

constexpr int kMmaWarps = 2;  // (3 warps x 6 CTAs/SM needs <= 112 registers: measured 4 % slower, small spills)

// S_CT > 0: S known at compile time (masks and fully-padded key tiles are pruned); 0: runtime S.
template <int S_PAD, int S_CT>
__global__ void __launch_bounds__(kMmaWarps * 32)
group_attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                           int S_rt, int H, float scale_log2e, int q_rows) {
  constexpr int MT = S_PAD / 16;  // query m-tiles (also 16-key steps)
  constexpr int NT = S_PAD / 8;   // 8-key n-tiles
  const int S = S_CT > 0 ? S_CT : S_rt;
  __shared__ __align__(128) uint8_t smem[3 * S_PAD * 128];
  const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const uint32_t sK = sQ + S_PAD * 128;
  const uint32_t sV = sK + S_PAD * 128;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int64_t prob = blockIdx.x;
  const int64_t g = prob / H;
  const int h = static_cast<int>(prob - g * H);
  const int D = H * kHeadDim;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const __nv_bfloat16* base = qkv + (g * S) * ld + h * kHeadDim;

  // ---- stage Q, K, V (S rows x 128 B each) with cp.async; zero the padding rows ----
  {
    constexpr int kRowsPerPass = kMmaWarps * 32 / 8;
    const int c = tid & 7;          // 16-byte chunk of the row (constant per thread)
    const int r_first = tid >> 3;
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      const uint32_t dst0 = sQ + which * (S_PAD * 128);
      const __nv_bfloat16* s0 = base + which * D + c * 8;
#pragma unroll
      for (int r = r_first; r < S_PAD; r += kRowsPerPass) {
        const uint32_t dst = dst0 + static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4));
        // Q rows are only needed up to the last computed 16-row query tile (q_rows = 1: 16 of 86 rows)
        if (r < S && (which != 0 || r < ((q_rows + 15) & ~15))) {
          cp_async_16(dst, s0 + static_cast<int64_t>(r) * ld);
        } else {
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
        }
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int gq = lane >> 2;  // fragment row group
  const int tq = lane & 3;   // fragment column pair
  // Per-lane ldmatrix offsets.  Row offsets added later are multiples of 8 rows, so the XOR
  // swizzle term (row & 7) == (lane & 7) is a per-lane constant.
  const uint32_t x7 = static_cast<uint32_t>(lane & 7);
  uint32_t q_off[4], k_off[4], v_off[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    q_off[i] = static_cast<uint32_t>(((lane & 7) + ((lane >> 3) & 1) * 8) * 128) + (((2 * i + (lane >> 4)) ^ x7) << 4);
    k_off[i] = static_cast<uint32_t>(((lane & 7) + ((lane >> 4) & 1) * 8) * 128) + (((2 * i + ((lane >> 3) & 1)) ^ x7) << 4);
    v_off[i] = q_off[i];  // V (transposed load) uses the same lane -> (row, chunk) pattern as Q
  }
  // number of 16-key steps / 8-key tiles that contain at least one real key
  const int kt_live = (S + 15) >> 4;

  for (int mt = warp; mt < MT; mt += kMmaWarps) {
    const int m0 = mt * 16;
    if (m0 >= q_rows) break;
    // ---- Q fragments (A operand), 4 k-steps of 16 ----
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldmatrix_x4(sQ + m0 * 128 + q_off[ks], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    // ---- scores = Q K^T ----
    float sc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
    for (int np = 0; np < MT; ++np) {
      if (S_CT > 0 ? (np * 16 < S_CT) : (np < kt_live)) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(sK + np * 2048 + k_off[ks], b0, b1, b2, b3);
          mma_bf16_16816(sc[2 * np], qa[ks], b0, b1);
          if (S_CT == 0 || np * 16 + 8 < S_CT) mma_bf16_16816(sc[2 * np + 1], qa[ks], b2, b3);
        }
      }
    }
    // ---- softmax over keys (rows gq and gq+8 of this m-tile) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (S_CT > 0 && nt * 8 + 8 <= S_CT) {
        // tile entirely inside the real keys: no masking
      } else {
        const int col = nt * 8 + 2 * tq;
        if (col >= S) sc[nt][0] = sc[nt][2] = -INFINITY;
        if (col + 1 >= S) sc[nt][1] = sc[nt][3] = -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float off0 = mx0 * scale_log2e, off1 = mx1 * scale_log2e;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (S_CT > 0 && nt * 8 >= S_CT) {  // tile entirely padding: probabilities are exactly 0
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      } else {
        sc[nt][0] = exp2f(fmaf(sc[nt][0], scale_log2e, -off0));
        sc[nt][1] = exp2f(fmaf(sc[nt][1], scale_log2e, -off0));
        sc[nt][2] = exp2f(fmaf(sc[nt][2], scale_log2e, -off1));
        sc[nt][3] = exp2f(fmaf(sc[nt][3], scale_log2e, -off1));
        sum0 += sc[nt][0] + sc[nt][1];
        sum1 += sc[nt][2] + sc[nt][3];
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // ---- O = P V ----
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < MT; ++kk) {
      if (S_CT > 0 ? (kk * 16 < S_CT) : (kk < kt_live)) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
        pa[1] = pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]);
        pa[2] = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(sV + kk * 2048 + v_off[dp], b0, b1, b2, b3);
          mma_bf16_16816(o[2 * dp], pa, b0, b1);
          mma_bf16_16816(o[2 * dp + 1], pa, b2, b3);
        }
      }
    }
    // ---- normalise and store (heads merged: column h*64 + d) ----
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    const int r0 = m0 + gq, r1 = m0 + gq + 8;
    __nv_bfloat16* obase = out + (g * q_rows) * D + h * kHeadDim + 2 * tq;
    const bool st0 = r0 < q_rows && r0 < S, st1 = r1 < q_rows && r1 < S;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (st0)
        *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(r0) * D + nt * 8) =
            pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
      if (st1)
        *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(r1) * D + nt * 8) =
            pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
}

template <int S_PAD, int S_CT>
int launch_mma(const void* qkv, void* out, int64_t groups, int S, int H, float scale, int q_rows,
               cudaStream_t st) {
  const int64_t problems = groups * H;
  if (problems >= (int64_t(1) << 31)) {
    set_error("duo_group_attention: too many groups");
    return DUO_ERR_INVALID;
  }
  group_attention_mma_kernel<S_PAD, S_CT><<<static_cast<unsigned>(problems), kMmaWarps * 32, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), S, H,
      scale * 1.4426950408889634f, q_rows);
  DUO_LAUNCH_CHECK("group_attention_mma_kernel");
  return DUO_OK;
}


// =============================== algo 4: small groups (S <= 8) ==============================
// The 2-scale model's scale attention: S = 6 tokens per patch.  One WARP per (patch, head) problem, eight problems
// per CTA pass, persistent grid-stride loop.  Everything lives in registers between two warp-level steps:
//   * the head slices of Q, K, V (S rows x 128 B each, 2.3 KB per problem) come in with coalesced 16-byte cp.async
//     into the warp's private 4 KB of shared memory (Q padded to 16 rows, K / V to 8 rows; the padding rows are
//     zeroed once and never written);
//   * scores = Q K^T as 4 x mma.sync.m16n8k16 (one 8-key tile), softmax on the accumulator fragment — the 6 scores
//     of a row sit in one quad (3 lanes x 2), so max and sum are two xor-shuffles each —, the probabilities become
//     the A fragment of O = P V (8 x m16n8k16 whose upper k half is zero) without leaving registers;
//   * the 6 x 64 output is transposed through the warp's (dead) Q rows and stored as complete 128-byte rows.
// ~120 instructions per problem (the generic FMA kernel: ~650, with 26 of 32 lanes idle in the score phase), so
// the kernel is bound by its 2.3 KB + 0.8 KB of HBM traffic per problem: every warp double-buffers its staging
// area (the next problem's cp.async loads fly while the current one is computed), 24 warps per SM.
constexpr int kSmallWarps = 8;
constexpr uint32_t kSmallBufBytes = 4096;   // Q: 16 rows | K: 8 rows | V: 8 rows, 128 B each
constexpr uint32_t kSmallWarpBytes = 2 * kSmallBufBytes;

__device__ __forceinline__ void mma_bf16_16816_lo(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  // A columns 8..15 and B rows 8..15 are zero (keys 8..15 do not exist)
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %6}, "
      "{%7, %6}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(0u), "r"(b0));
}

template <int S_CT>
__global__ void __launch_bounds__(kSmallWarps * 32, 3)
group_attention_small_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                             int64_t problems, int S_rt, int H, float scale_log2e, int q_rows) {
  const int S = S_CT > 0 ? S_CT : S_rt;
  extern __shared__ __align__(128) uint8_t smem_small[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t s_warp = static_cast<uint32_t>(__cvta_generic_to_shared(smem_small)) + static_cast<uint32_t>(warp) * kSmallWarpBytes;
  for (uint32_t off = static_cast<uint32_t>(lane) * 16; off < kSmallWarpBytes; off += 32 * 16)
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(s_warp + off), "r"(0u) : "memory");
  __syncwarp();

  const int D = H * kHeadDim;
  const int64_t ld = 3 * static_cast<int64_t>(D);
  const int gq = lane >> 2;  // fragment row (query) / B column (key or output dim)
  const int tq = lane & 3;   // fragment column pair
  const uint32_t x7 = static_cast<uint32_t>(lane & 7);
  // ldmatrix lane addresses (row & 7 == lane & 7 for all of them)
  uint32_t q_off[4], k_off[2], v_off[2];
#pragma unroll
  for (int i = 0; i < 4; ++i)  // Q, k-step i: matrices (rows 0-7 | 8-15) x (chunk 2i | 2i+1)
    q_off[i] = static_cast<uint32_t>(((lane & 7) + ((lane >> 3) & 1) * 8) * 128) + (((2 * i + (lane >> 4)) ^ x7) << 4);
#pragma unroll
  for (int i = 0; i < 2; ++i) {  // K: chunks 4i .. 4i+3 of keys 0-7 (k-steps 2i, 2i+1); V (transposed): the same
    k_off[i] = static_cast<uint32_t>((lane & 7) * 128) + (((4 * i + (lane >> 3)) ^ x7) << 4);
    v_off[i] = k_off[i];
  }
  const int stage_chunks = 24 * S;  // 3 slices x S rows x 8 chunks of 16 bytes
  const int out_chunks = 8 * q_rows;
  auto stage = [&](int64_t pr, uint32_t buf) {  // cp.async the Q | K | V head slices of problem pr into buffer buf
    const int64_t pg = pr / H;
    const int ph = static_cast<int>(pr - pg * H);
    const __nv_bfloat16* src = qkv + (pg * S) * ld + ph * kHeadDim;
    for (int idx = lane; idx < stage_chunks; idx += 32) {
      const int which = idx / (8 * S);
      const int rem = idx - which * 8 * S;
      const int r = rem >> 3, c = rem & 7;
      const uint32_t dst = buf + (which == 0 ? 0u : (which == 1 ? 16u * 128u : 24u * 128u)) + swz(r, c);
      cp_async_16(dst, src + which * D + static_cast<int64_t>(r) * ld + c * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t first = static_cast<int64_t>(blockIdx.x) * kSmallWarps + warp;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kSmallWarps;
  if (first < problems) stage(first, s_warp);
  uint32_t cur = 0;

  for (int64_t prob = first; prob < problems; prob += stride, cur ^= 1u) {
    const int64_t g = prob / H;
    const int h = static_cast<int>(prob - g * H);
    const uint32_t sQ = s_warp + cur * kSmallBufBytes;
    const uint32_t sK = sQ + 16 * 128;
    const uint32_t sV = sK + 8 * 128;
    if (prob + stride < problems) {  // the other buffer's previous user finished with the __syncwarp below
      stage(prob + stride, s_warp + (cur ^ 1u) * kSmallBufBytes);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();

    // ---- scores = Q K^T: one 8-key tile, 4 k-steps ----
    float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      uint32_t kb0, kb1, kb2, kb3;  // (k-step 2kp: b0, b1), (k-step 2kp+1: b0, b1)
      ldmatrix_x4(sK + k_off[kp], kb0, kb1, kb2, kb3);
      uint32_t qa[4];
      ldmatrix_x4(sQ + q_off[2 * kp], qa[0], qa[1], qa[2], qa[3]);
      mma_bf16_16816(sc, qa, kb0, kb1);
      ldmatrix_x4(sQ + q_off[2 * kp + 1], qa[0], qa[1], qa[2], qa[3]);
      mma_bf16_16816(sc, qa, kb2, kb3);
    }
    // ---- softmax over the keys of row gq (sc[0], sc[1]); rows 8..15 (sc[2], sc[3]) are padding queries ----
    const int col = 2 * tq;
    float s0 = col < S ? sc[0] : -INFINITY;
    float s1 = col + 1 < S ? sc[1] : -INFINITY;
    float mx = fmaxf(s0, s1);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float off = mx * scale_log2e;
    const float p0 = col < S ? ex2_approx(fmaf(s0, scale_log2e, -off)) : 0.f;
    const float p1 = col + 1 < S ? ex2_approx(fmaf(s1, scale_log2e, -off)) : 0.f;
    float sum = p0 + p1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    const uint32_t pa0 = pack_bf16x2(p0, p1);  // A fragment of P: row gq, keys 2tq, 2tq+1 (rows 8..15: unused)

    // ---- O = P V: 8 output tiles of 8 dims, keys 0..7 only ----
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int dq = 0; dq < 2; ++dq) {
      uint32_t vb0, vb1, vb2, vb3;  // keys 0-7 x dims of chunks 4dq .. 4dq+3, transposed
      ldmatrix_x4_trans(sV + v_off[dq], vb0, vb1, vb2, vb3);
      mma_bf16_16816_lo(o[4 * dq + 0], pa0, 0u, vb0);
      mma_bf16_16816_lo(o[4 * dq + 1], pa0, 0u, vb1);
      mma_bf16_16816_lo(o[4 * dq + 2], pa0, 0u, vb2);
      mma_bf16_16816_lo(o[4 * dq + 3], pa0, 0u, vb3);
    }
    // ---- row gq of O -> the warp's Q row gq (dead), then complete 128-byte rows to global ----
    __syncwarp();
    if (gq < S) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + swz(gq, nt) + 4u * static_cast<uint32_t>(tq)),
                     "r"(pack_bf16x2(o[nt][0] * inv, o[nt][1] * inv))
                     : "memory");
    }
    __syncwarp();
    __nv_bfloat16* obase = out + (g * q_rows) * D + h * kHeadDim;
    for (int idx = lane; idx < out_chunks; idx += 32) {
      const int r = idx >> 3, c = idx & 7;
      uint32_t w0, w1, w2, w3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(sQ + swz(r, c)));
      *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(r) * D + c * 8) = make_uint4(w0, w1, w2, w3);
    }
    __syncwarp();  // the staging rows are free for the next problem's loads
  }
}

int launch_small(const void* qkv, void* out, int64_t groups, int S, int H, float scale, int q_rows, cudaStream_t st) {
  const int64_t problems = groups * H;
  const int64_t ctas_needed = (problems + kSmallWarps - 1) / kSmallWarps;
  const int64_t ctas_max = 3LL * device_sm_count();
  const unsigned grid = static_cast<unsigned>(ctas_needed < ctas_max ? ctas_needed : ctas_max);
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  constexpr int kSmem = kSmallWarps * kSmallWarpBytes;  // 64 KB: three CTAs per SM
  static uint64_t configured = 0;  // per device
  if (first_use_on_device(configured)) {
    DUO_CUDA(cudaFuncSetAttribute(group_attention_small_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    DUO_CUDA(cudaFuncSetAttribute(group_attention_small_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  if (S == 6)
    group_attention_small_kernel<6><<<grid, kSmallWarps * 32, kSmem, st>>>(q, o, problems, S, H, scale * 1.4426950408889634f, q_rows);
  else
    group_attention_small_kernel<0><<<grid, kSmallWarps * 32, kSmem, st>>>(q, o, problems, S, H, scale * 1.4426950408889634f, q_rows);
  DUO_LAUNCH_CHECK("group_attention_small_kernel");
  return DUO_OK;
}

}  // namespace
}  // namespace duo

namespace duo {
int launch_scale_attention_tc(const void* qkv, void* out, int64_t groups, int S, int H, float scale, cudaStream_t st);
int launch_patch_attention_tc(const void* qkv, void* out, int64_t groups, int N, int H, float scale, cudaStream_t st);
}

extern "C" int duo_group_attention(const void* qkv, int32_t in_kind, void* out, int32_t out_kind,
                                   int64_t num_groups, int32_t S, int32_t num_heads, float scale,
                                   int32_t algo, int32_t q_rows, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(qkv && out, "duo_group_attention: NULL pointer");
  DUO_CHECK_ARG(num_groups > 0 && S > 0 && num_heads > 0, "duo_group_attention: empty problem");
  DUO_CHECK_ARG(in_kind == DUO_ACT_BF16 || in_kind == DUO_ACT_F32 || in_kind == DUO_ACT_SPLIT,
                "duo_group_attention: in_kind=%d", in_kind);
  DUO_CHECK_ARG(out_kind == DUO_ACT_BF16 || out_kind == DUO_ACT_SPLIT || out_kind == DUO_ACT_F32,
                "duo_group_attention: out_kind=%d", out_kind);
  DUO_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "duo_group_attention: pointers must be 16-byte aligned");
  DUO_CHECK_ARG(q_rows >= 1 && q_rows <= S, "duo_group_attention: q_rows=%d must be in [1, S=%d]", q_rows, S);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (in_kind == DUO_ACT_SPLIT) {  // hi | lo bf16 pairs in and out: the split-precision tcgen05 kernel only
    DUO_CHECK_ARG(out_kind == DUO_ACT_SPLIT && S <= 64 && q_rows == S && (algo == 0 || algo == 3),
                  "duo_group_attention: split input needs split output, S <= 64, q_rows == S (S=%d q_rows=%d algo=%d)", S,
                  q_rows, algo);
    return launch_patch_attention_tc(qkv, out, num_groups, S, num_heads, scale, st);
  }
  const bool mma_ok = in_kind == DUO_ACT_BF16 && out_kind == DUO_ACT_BF16 && S > 16 && S <= 96;
  const bool small_ok = in_kind == DUO_ACT_BF16 && out_kind == DUO_ACT_BF16 && S <= 8;
  // auto: tcgen05 kernel for the 4-scale group size, mma.sync kernel for the other tensor-core sizes, the
  // warp-per-(patch, head) register kernel for the 2-scale group size
  if (algo == 0) algo = (mma_ok && S > 64 && q_rows == S) ? 3 : (mma_ok ? 2 : (small_ok ? 4 : 1));
  if (algo == 4) {
    DUO_CHECK_ARG(small_ok, "duo_group_attention: algo 4 needs bf16 in/out and S <= 8 (S=%d)", S);
    return launch_small(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
  }
  if (algo == 3) {  // tcgen05 / TMEM kernel (scale_attention_tc.cu)
    DUO_CHECK_ARG(mma_ok && S > 64 && q_rows == S,
                  "duo_group_attention: algo 3 needs bf16 in/out, 64 < S <= 96 and q_rows == S (S=%d q_rows=%d)", S, q_rows);
    return launch_scale_attention_tc(qkv, out, num_groups, S, num_heads, scale, st);
  }
  if (algo == 2) {
    DUO_CHECK_ARG(mma_ok, "duo_group_attention: algo 2 needs bf16 in/out and 16 < S <= 96 (S=%d)", S);
    // the model's own sizes get compile-time S (masks / padded tiles pruned)
    if (S == 86) return launch_mma<96, 86>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    if (S == 22) return launch_mma<32, 22>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    if (S == 50) return launch_mma<64, 50>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    if (S <= 32) return launch_mma<32, 0>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    if (S <= 48) return launch_mma<48, 0>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    if (S <= 64) return launch_mma<64, 0>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    if (S <= 80) return launch_mma<80, 0>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    return launch_mma<96, 0>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
  }
  DUO_CHECK_ARG(algo == 1, "duo_group_attention: algo=%d", algo);
  if (in_kind == DUO_ACT_BF16) {
    switch (out_kind) {
      case DUO_ACT_BF16:
        return dispatch_fma_kpl<__nv_bfloat16, DUO_ACT_BF16>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
      case DUO_ACT_SPLIT:
        return dispatch_fma_kpl<__nv_bfloat16, DUO_ACT_SPLIT>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
      default:
        return dispatch_fma_kpl<__nv_bfloat16, DUO_ACT_F32>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    }
  }
  switch (out_kind) {
    case DUO_ACT_BF16:
      return dispatch_fma_kpl<float, DUO_ACT_BF16>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    case DUO_ACT_SPLIT:
      return dispatch_fma_kpl<float, DUO_ACT_SPLIT>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
    default:
      return dispatch_fma_kpl<float, DUO_ACT_F32>(qkv, out, num_groups, S, num_heads, scale, q_rows, st);
  }
}
