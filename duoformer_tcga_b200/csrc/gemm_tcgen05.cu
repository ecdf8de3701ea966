// Persistent, warp-specialised TMA + tcgen05/TMEM GEMM for sm_100a.
//
//   C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue), bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every nn.Linear / 1x1 Conv2d of the reference forward (see include/duoformer_sm100.h
// for the file:line map).  Two kernels share one epilogue:
//
//   gemm_tcgen05_pair_kernel  (large problems) a cluster of two CTAs on one SM pair computes a
//       256 x 256 tile with tcgen05.mma.cta_group::2 — see the block comment above that kernel;
//   gemm_tcgen05_kernel       (small problems) one CTA per 128 x {128,256} tile.
//
// Common structure (persistent CTAs, static round-robin tile scheduler, N fastest so concurrently
// running CTAs share one A row-panel in L2):
//
//   epilogue warps 0..3 (0..7 for the GELU epilogue of the pair kernel): tcgen05.ld the fp32
//               accumulator (one TMEM lane quarter per warp, one output row per thread), fuse
//               bias / GELU(erf) / LayerScale, then
//                 * bf16 outputs: rows staged in a per-warp 128B-swizzled shared-memory tile and
//                   written with TMA stores (full-line, coalesced);
//                 * residual (X += ...): staged fp32 tile + TMA reduce-add into the fp32 residual
//                   stream (the read-modify-write happens in L2, no SM-side loads);
//                 * token scatter: rows transposed through the staging tile, four complete
//                   128-byte lines per store instruction;
//                 * fp32 / hi-lo split outputs (fp32 mode only): direct 16-byte global stores.
//               Double-buffered TMEM lets the epilogue of tile i overlap the MMAs of tile i+1.
//   TMA warp    one lane streams A (128x64) and W (BLOCK_N x 64) tiles (K-major, SWIZZLE_128B)
//               into a kStages-deep shared-memory ring, completion on `full` mbarriers.
//   MMA warp    one lane issues tcgen05.mma (K = 16 per instruction) into one of two TMEM
//               accumulator buffers; tcgen05.commit releases ring slots (`empty`) and publishes
//               finished accumulators (`tmem_full`).
//
// split3 mode (fp32-accuracy path): A and W hold bf16 hi|lo halves; the K loop runs three
// segments (Ah*Wh, Ah*Wl, Al*Wh) into the same accumulator.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace duo {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumThreads = 192;
constexpr int kNumEpilogueThreads = 128;
constexpr uint32_t kStagingBytesPerWarp = 2 * 32 * 128;  // two 32-row x 128 B buffers

// Internal epilogue variants (superset of the ABI's DUO_EPI_*): staged TMA paths.
constexpr int kEpiResidualTma = 100;  // DUO_EPI_RESIDUAL_F32 through TMA reduce-add
constexpr int kEpiResidualLn = 101;   // residual update + fused LayerNorm of the updated rows (pair kernel)

template <int EPI>
struct EpiTraits {
  static constexpr bool kStagedBf16 = (EPI == DUO_EPI_BF16 || EPI == DUO_EPI_GELU_BF16);
  static constexpr bool kStagedF32 = (EPI == kEpiResidualTma || EPI == kEpiResidualLn);
  static constexpr bool kStaged = kStagedBf16 || kStagedF32;
};

template <int BLOCK_N>
struct Cfg {
  static constexpr int kStages = BLOCK_N == 256 ? 4 : 6;
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;   // 16 KB
  static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;   // 32 / 16 KB
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;           // two accumulator buffers
  static constexpr uint32_t kStagingBytes = 4 * kStagingBytesPerWarp;  // 32 KB
  static constexpr uint32_t kBarrierBytes = (2 * kStages + 4) * 8 + 8;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024;
};

struct GemmParams {
  const float* bias;
  void* out;
  const float* gamma;
  const float* ln_gamma;  // fused LayerNorm (kEpiResidualLn)
  const float* ln_beta;
  void* ln_out;
  uint32_t* ln_sync;  // [num_m_blocks][2 CTAs][4 quarters] finished N tiles per 32-row slab
  float ln_eps;
  int32_t relu;           // BF16 / F32 epilogues: clamp at zero
  const int32_t* row_map;
  const float* pos;
  int64_t M;
  int64_t ldo;
  int32_t N, K;
  int32_t split3;
  int32_t rows_per_group, dest_rows_per_group, pos_period;
  int32_t num_m_blocks, num_n_blocks;
  uint32_t idesc_mask;  // ~0, or with the a_format / b_format bits cleared (fp16 operands instead of bf16)
};

// acc[32] (fp32 bits) -> f[32] = acc + bias (optionally GELU'd / scaled by LayerScale gamma)
// Bias slice [col, col+32) -> registers; issued BEFORE waiting on the TMEM load so both latencies overlap.
__device__ __forceinline__ void epilogue_bias_load(const GemmParams& p, int col, float4 (&b)[8]) {
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = __ldg(b4 + j);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_math(const GemmParams& p, int col, const uint32_t (&v)[32],
                                              const float4 (&b)[8], float (&f)[32]) {
  if constexpr (EPI == DUO_EPI_GELU_BF16) {  // bias add and GELU on packed fp32 pairs
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint64_t lo = add2(pack2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack2(b[j].x, b[j].y));
      const uint64_t hi = add2(pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack2(b[j].z, b[j].w));
      unpack2(gelu_erf_sigmoid_p2(lo), f[4 * j + 0], f[4 * j + 1]);
      unpack2(gelu_erf_sigmoid_p2(hi), f[4 * j + 2], f[4 * j + 3]);
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b[j].x;
    f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b[j].y;
    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b[j].z;
    f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b[j].w;
  }
  if constexpr (EPI == DUO_EPI_BF16 || EPI == DUO_EPI_F32) {
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
  }
  if constexpr (EPI == DUO_EPI_GELU_SPLIT_BF16) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
  }
  if constexpr (EPI == kEpiResidualTma) {
    if (p.gamma != nullptr) {
      const float4* g4 = reinterpret_cast<const float4*>(p.gamma + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g = __ldg(g4 + j);
        f[4 * j + 0] *= g.x;
        f[4 * j + 1] *= g.y;
        f[4 * j + 2] *= g.z;
        f[4 * j + 3] *= g.w;
      }
    }
  }
}

// Direct (non-staged) stores: one output row per thread, 32 consecutive columns.
template <int EPI>
__device__ __forceinline__ void epilogue_store_direct(const GemmParams& p, int64_t row, int col,
                                                      float (&f)[32]) {
  if constexpr (EPI == DUO_EPI_SPLIT_BF16 || EPI == DUO_EPI_GELU_SPLIT_BF16) {
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + col;
    __nv_bfloat16* ol = oh + p.N;
    uint4* h4 = reinterpret_cast<uint4*>(oh);
    uint4* l4 = reinterpret_cast<uint4*>(ol);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 h, l;
      pack_split2(f[8 * j + 0], f[8 * j + 1], h.x, l.x);
      pack_split2(f[8 * j + 2], f[8 * j + 3], h.y, l.y);
      pack_split2(f[8 * j + 4], f[8 * j + 5], h.z, l.z);
      pack_split2(f[8 * j + 6], f[8 * j + 7], h.w, l.w);
      h4[j] = h;
      l4[j] = l;
    }
  } else if constexpr (EPI == DUO_EPI_F32) {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + col);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  } else if constexpr (EPI == DUO_EPI_RESIDUAL_F32) {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + col);
    float4 r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = o4[j];
    if (p.gamma != nullptr) {
      const float4* g4 = reinterpret_cast<const float4*>(p.gamma + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g = __ldg(g4 + j);
        r[j].x = fmaf(g.x, f[4 * j + 0], r[j].x);
        r[j].y = fmaf(g.y, f[4 * j + 1], r[j].y);
        r[j].z = fmaf(g.z, f[4 * j + 2], r[j].z);
        r[j].w = fmaf(g.w, f[4 * j + 3], r[j].w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r[j].x += f[4 * j + 0];
        r[j].y += f[4 * j + 1];
        r[j].z += f[4 * j + 2];
        r[j].w += f[4 * j + 3];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o4[j] = r[j];
  } else if constexpr (EPI == DUO_EPI_SCATTER_F32) {
    const int64_t grp = row / p.rows_per_group;
    const int32_t in_grp = static_cast<int32_t>(row - grp * p.rows_per_group);
    const int32_t dst_in_grp = __ldg(p.row_map + in_grp);
    const int64_t dst = grp * p.dest_rows_per_group + dst_in_grp;
    if (p.pos != nullptr) {
      const int32_t s = dst_in_grp % p.pos_period;
      const float4* q4 = reinterpret_cast<const float4*>(p.pos + static_cast<int64_t>(s) * p.N + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = __ldg(q4 + j);
        f[4 * j + 0] += q.x;
        f[4 * j + 1] += q.y;
        f[4 * j + 2] += q.z;
        f[4 * j + 3] += q.w;
      }
    }
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + dst * p.ldo + col);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// One accumulator slab (this warp's 32 rows x columns [c_begin, c_end) of the tile): TMEM ->
// registers -> fused math -> global memory.  `release()` is called as soon as this warp has read
// its part of the accumulator completely.
template <int EPI, int NBUF, typename ReleaseFn>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const CUtensorMap* tmap_out,
                                              uint32_t taddr, int row0, int lane, int n0, int c_begin,
                                              int c_end, uint32_t stg, uint32_t& stg_buf,
                                              ReleaseFn release) {
  using ET = EpiTraits<EPI>;
  const int64_t row = static_cast<int64_t>(row0) + lane;
  const bool valid = row < p.M;
  const uint32_t my_row_off = static_cast<uint32_t>(lane) * 128u;
  (void)valid;
  (void)my_row_off;
  if constexpr (ET::kStagedBf16) {
    // 64 output columns (= 128 B of bf16) per staged chunk, filled 32 columns at a time
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 64) {
      const uint32_t buf = stg + stg_buf * (32u * 128u) + my_row_off;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        float4 bia[8];
        ptx::tmem_ld_32x32(taddr + static_cast<uint32_t>(c + 32 * h), v);
        epilogue_bias_load(p, n0 + c + 32 * h, bia);
        ptx::tmem_ld_wait();
        if (h == 1 && c + 64 >= c_end) {  // accumulator fully read: hand the TMEM buffer back early
          ptx::tc_fence_before();
          release();
        }
        float f[32];
        epilogue_math<EPI>(p, n0 + c + 32 * h, v, bia, f);
        if (h == 0) {
          if (lane == 0) ptx::tma_store_wait_read<NBUF - 1>();  // buffer `stg_buf` no longer being read
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)  // 16-byte chunk (4h + j) of this row, XOR-swizzled
          st_shared_v4(buf + (static_cast<uint32_t>((4 * h + j) ^ (lane & 7)) << 4),
                       pack_bf16x2(f[8 * j + 0], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                       pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(tmap_out, stg + stg_buf * (32u * 128u), n0 + c, row0);
        ptx::tma_store_commit();
      }
      stg_buf = (NBUF == 1) ? 0u : (stg_buf ^ 1u);
    }
  } else if constexpr (ET::kStagedF32) {
    // 32 output columns (= 128 B of fp32) per staged chunk, TMA reduce-add into X
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t v[32];
      float4 bia[8];
      ptx::tmem_ld_32x32(taddr + static_cast<uint32_t>(c), v);
      epilogue_bias_load(p, n0 + c, bia);
      ptx::tmem_ld_wait();
      if (c + 32 >= c_end) {
        ptx::tc_fence_before();
        release();
      }
      float f[32];
      epilogue_math<EPI>(p, n0 + c, v, bia, f);
      if (lane == 0) ptx::tma_store_wait_read<NBUF - 1>();
      __syncwarp();
      const uint32_t buf = stg + stg_buf * (32u * 128u) + my_row_off;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        st_shared_v4(buf + (static_cast<uint32_t>(j ^ (lane & 7)) << 4), __float_as_uint(f[4 * j + 0]),
                     __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                     __float_as_uint(f[4 * j + 3]));
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_reduce_add_2d(tmap_out, stg + stg_buf * (32u * 128u), n0 + c, row0);
        ptx::tma_store_commit();
      }
      stg_buf = (NBUF == 1) ? 0u : (stg_buf ^ 1u);
    }
  } else if constexpr (EPI == DUO_EPI_SCATTER_F32) {
    // Token scatter: every source row goes to its own destination row (p*S + s), so no tensor
    // store applies.  Rows are transposed through the warp's swizzled staging tile so that each
    // store instruction writes four complete 128-byte lines (8 lanes x 16 B per row) instead of
    // 32 partial ones.
    int64_t dst_row = -1;
    int32_t dst_s = 0;
    if (valid) {
      const int64_t grp = row / p.rows_per_group;
      const int32_t in_grp = static_cast<int32_t>(row - grp * p.rows_per_group);
      const int32_t dst_in_grp = __ldg(p.row_map + in_grp);
      dst_row = grp * p.dest_rows_per_group + dst_in_grp;
      dst_s = p.pos != nullptr ? dst_in_grp % p.pos_period : 0;
    }
    const int sub_row = lane >> 3;  // row inside a group of four handled by one store instruction
    const int chunk = lane & 7;     // 16-byte chunk of the 128-byte row
    const uint32_t buf0 = stg + stg_buf * (32u * 128u);
    // after the transpose lane (sub_row, chunk) stores rows 4i + sub_row, i = 0..7: their destination
    // and positional rows are fixed for the whole tile
    float* orow[8];
    const float* prow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + sub_row;
      const int64_t d = __shfl_sync(0xffffffffu, dst_row, r);
      const int32_t ds = __shfl_sync(0xffffffffu, dst_s, r);
      orow[i] = d >= 0 ? reinterpret_cast<float*>(p.out) + d * p.ldo + n0 + 4 * chunk : nullptr;
      prow[i] = p.pos != nullptr ? p.pos + static_cast<int64_t>(ds) * p.N + n0 + 4 * chunk : nullptr;
    }
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t v[32];
      float4 bia[8];
      float4 q[8];
      ptx::tmem_ld_32x32(taddr + static_cast<uint32_t>(c), v);
      epilogue_bias_load(p, n0 + c, bia);
#pragma unroll
      for (int i = 0; i < 8; ++i)  // positional slices issued together, ahead of the transpose
        q[i] = prow[i] != nullptr ? __ldg(reinterpret_cast<const float4*>(prow[i] + c))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      ptx::tmem_ld_wait();
      if (c + 32 >= c_end) {
        ptx::tc_fence_before();
        release();
      }
      float f[32];
      epilogue_math<EPI>(p, n0 + c, v, bia, f);
      __syncwarp();  // previous chunk's reads of the staging tile are done
#pragma unroll
      for (int j = 0; j < 8; ++j)
        st_shared_v4(buf0 + my_row_off + (static_cast<uint32_t>(j ^ (lane & 7)) << 4),
                     __float_as_uint(f[4 * j + 0]), __float_as_uint(f[4 * j + 1]),
                     __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + sub_row;  // row of the warp slab this lane now stores
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                     : "r"(buf0 + static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(chunk ^ (r & 7)) << 4)));
        if (orow[i] != nullptr)
          *reinterpret_cast<float4*>(orow[i] + c) =
              make_float4(__uint_as_float(w0) + q[i].x, __uint_as_float(w1) + q[i].y, __uint_as_float(w2) + q[i].z,
                          __uint_as_float(w3) + q[i].w);
      }
    }
  } else {
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t v[32];
      float4 bia[8];
      ptx::tmem_ld_32x32(taddr + static_cast<uint32_t>(c), v);
      epilogue_bias_load(p, n0 + c, bia);
      ptx::tmem_ld_wait();
      if (c + 32 >= c_end) {
        ptx::tc_fence_before();
        release();
      }
      if (valid) {
        float f[32];
        epilogue_math<EPI>(p, n0 + c, v, bia, f);
        epilogue_store_direct<EPI>(p, row, n0 + c, f);
      }
    }
  }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
  using C = Cfg<BLOCK_N>;
  using ET = EpiTraits<EPI>;
  constexpr int kStages = C::kStages;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment.
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + kStages * C::kStageBytes;  // 1024-aligned
  const uint32_t bar_base = staging_base + C::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  uint32_t* tmem_ptr_generic =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_ptr_smem - ptx::smem_u32(smem_raw)));

  // Warp roles: epilogue warps 0..3, then the TMA producer and the MMA issuer as the HIGHEST warp
  // ids — the SM's warp arbiter favours higher warp ids, and the two single-thread roles must
  // never be starved of issue slots by epilogue math.
  constexpr int kTmaWarp = 4, kMmaWarp = 5;
  const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == kTmaWarp && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if constexpr (ET::kStaged) ptx::prefetch_tmap(&tmap_out);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar(a), 1);
      ptx::mbar_init(tmem_empty_bar(a), kNumEpilogueThreads / 32);  // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == kMmaWarp) {
    ptx::tmem_alloc<C::kTmemCols>(tmem_ptr_smem);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  const int kseg_blocks = p.K / kBlockK;
  const int num_k_blocks = p.split3 == 1 ? 3 * kseg_blocks : (p.split3 == 2 ? 2 * kseg_blocks : kseg_blocks);
  const int64_t num_tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;

  if (warp_idx == kTmaWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = static_cast<int>(tile / p.num_n_blocks);
        const int n_blk = static_cast<int>(tile - static_cast<int64_t>(m_blk) * p.num_n_blocks);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          int a_k, b_k;
          if (p.split3 == 1) {  // Ah*Wh, Ah*Wl, Al*Wh
            const int seg = kb / kseg_blocks;
            const int r = kb - seg * kseg_blocks;
            a_k = (seg == 2 ? p.K : 0) + r * kBlockK;
            b_k = (seg == 1 ? p.K : 0) + r * kBlockK;
          } else if (p.split3 == 2) {  // A is plain bf16: A*Wh, A*Wl
            const int seg = kb / kseg_blocks;
            const int r = kb - seg * kseg_blocks;
            a_k = r * kBlockK;
            b_k = seg * p.K + r * kBlockK;
          } else {
            a_k = b_k = kb * kBlockK;
          }
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
          ptx::mbar_arrive_expect_tx(full_bar(stage), C::kStageBytes);
          ptx::tma_load_2d(sa, &tmap_a, full_bar(stage), a_k, m_blk * kBlockM);
          ptx::tma_load_2d(sb, &tmap_b, full_bar(stage), b_k, n_blk * BLOCK_N);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(kBlockM, BLOCK_N) & p.idesc_mask;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
          const uint64_t desc_a = ptx::make_smem_desc_sw128(sa);
          const uint64_t desc_b = ptx::make_smem_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) address field
            ptx::umma_bf16(tmem_d, desc_a + static_cast<uint64_t>(2 * k),
                           desc_b + static_cast<uint64_t>(2 * k), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));  // ring slot free once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit(tmem_full_bar(acc));  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue warps (0..3) =====================
    const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
    const uint32_t stg = staging_base + static_cast<uint32_t>(quarter) * kStagingBytesPerWarp;
    uint32_t stg_buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = static_cast<int>(tile / p.num_n_blocks);
      const int n_blk = static_cast<int>(tile - static_cast<int64_t>(m_blk) * p.num_n_blocks);
      const int row0 = m_blk * kBlockM + quarter * 32;  // first row of this warp's slab
      const int n0 = n_blk * BLOCK_N;
      ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);

      epilogue_tile<EPI, 2>(p, &tmap_out, taddr, row0, lane, n0, 0, BLOCK_N, stg, stg_buf,
                         [&]() {
                           __syncwarp();
                           if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(acc));
                         });
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if constexpr (ET::kStaged) {
      if (lane == 0) ptx::tma_store_wait<0>();  // all bulk stores of this warp complete
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ===========================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs (one SM pair) computes a 256 x 256 tile.
// Each CTA loads its own 128 A rows and HALF of the W tile (128 of the 256 N rows) per k-block
// (32 KB / stage instead of 48 KB -> a 6-stage ring, one third less L2->SM operand traffic);
// the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16), which
// reads A from each CTA's shared memory, the two W halves from both, and accumulates rows
// [0,128) in the leader's TMEM and rows [128,256) in the peer's.  tcgen05.commit multicasts
// the "slot free" / "accumulator full" arrivals to both CTAs; each CTA runs its own epilogue.
// ===========================================================================================
constexpr int kPairBlockN = 256;
// EPI_WARPS epilogue warps: 4 (one per TMEM lane quarter, all 256 columns, 6 operand stages) or 8
// (two per quarter, 128 columns each, 5 operand stages — used when the epilogue is heavy: GELU,
// token scatter; both only occur with short K).  Staging is double-buffered per warp either way.
// FUSED_LN: four extra LayerNorm warps (one per epilogue warp) and a panel counter each.
constexpr int kLnWarps = 4;  // one per epilogue warp (32 rows each, four in flight)
template <int EPI_WARPS, bool FUSED_LN = false>
struct PairCfg {
  static constexpr int kStages = EPI_WARPS == 8 ? 5 : 6;
  static constexpr int kThreads = 64 + 32 * EPI_WARPS + (FUSED_LN ? 32 * kLnWarps : 0);
  static constexpr int kStagingBufs = 2;
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;          // 16 KB (this CTA's 128 rows)
  static constexpr uint32_t kBBytes = (kPairBlockN / 2) * kBlockK * 2;  // 16 KB (this CTA's N half)
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * kPairBlockN;
  static constexpr uint32_t kStagingBytes = EPI_WARPS * kStagingBufs * 32 * 128;  // 32 KB
  static constexpr uint32_t kBarrierBytes = (2 * kStages + 4) * 8 + 8 + (FUSED_LN ? kLnWarps * 4 : 0);
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024;
};

// ---- residual update + fused LayerNorm (kEpiResidualLn, pair kernel, row-panel tile order) ------
// The residual update is the ordinary TMA reduce-add epilogue in the ordinary tile order (the N
// tiles of a 256-row panel run on neighbouring pairs at the same time, which keeps the A panel
// in L2).  When an epilogue warp knows the reduce-adds of a tile have completed
// (cp.async.bulk.wait_group, checked one tile late so the store pipeline never drains) it bumps
// the global counter of its 32-row slab, p.ln_sync[m_blk][cta][quarter].  The pair that owns the
// panel's LAST N tile runs the LayerNorm: its LN warp waits until the slab's counter reaches the
// number of N tiles, resets it, re-reads the finished rows (L2 hits: they were just written
// there), normalises them exactly like layernorm.cu (two-pass statistics in registers) and
// writes the bf16 operand of the next GEMM.  LN warps never touch TMEM or the smem ring, so the
// GEMM pipeline does not wait on them; every CTA of the persistent grid is resident, so the
// producers of a counter are always running (a bounded spin traps instead of hanging).
__device__ __forceinline__ float4 ld_l2_f4(const float4* ptr) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(ptr)
               : "memory");
  return r;
}

template <int NV>
__device__ __forceinline__ void ln_row_store(const float4 (&v)[NV], const float4* g4, const float4* b4,
                                             float eps, int lane, __nv_bfloat16* yrow) {
  constexpr int D = NV * 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
    q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i);
    const float4 b = __ldg(b4 + lane + 32 * i);
    uint2 w;
    w.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
    w.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    reinterpret_cast<uint2*>(yrow)[lane + 32 * i] = w;
  }
}

// rows [row_begin, row_begin + nrows) of X (= p.out, fp32, leading dim p.ldo) -> p.ln_out (bf16 [M, N]);
// four rows in flight per warp (the loads are L2 / HBM latency bound).
template <int NV>
__device__ __forceinline__ void ln_rows_from_l2(const GemmParams& p, int64_t row_begin, int nrows, int lane) {
  constexpr int D = NV * 128;
  constexpr int kRows = NV > 6 ? 2 : 4;
  const float4* g4 = reinterpret_cast<const float4*>(p.ln_gamma);
  const float4* b4 = reinterpret_cast<const float4*>(p.ln_beta);
  const float* x0 = reinterpret_cast<const float*>(p.out) + row_begin * p.ldo;
  __nv_bfloat16* y0 = reinterpret_cast<__nv_bfloat16*>(p.ln_out) + row_begin * D;
#pragma unroll 1
  for (int r = 0; r < nrows; r += kRows) {
    float4 v[kRows][NV];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int rk = (r + k < nrows) ? r + k : nrows - 1;
      const float4* xr = reinterpret_cast<const float4*>(x0 + static_cast<int64_t>(rk) * p.ldo);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[k][i] = ld_l2_f4(xr + lane + 32 * i);
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k)
      if (r + k < nrows) ln_row_store<NV>(v[k], g4, b4, p.ln_eps, lane, y0 + static_cast<int64_t>(r + k) * D);
  }
}

// Tile order of the pair kernels: tiles round-robin over pairs, N fastest — the N tiles of a row panel
// run on neighbouring pairs at the same time and share the A panel through L2.  (Pair-owned row
// panels were measured: fc2 0.92 -> 1.16 ms per 64 images, 74 x 1.5 MB of A panels do not fit L2.)
__device__ __forceinline__ bool pair_tile(int64_t it, int64_t pair_idx, int64_t pair_stride, int nmb, int nnb,
                                          int& m_blk, int& n_blk) {
  const int64_t tile = pair_idx + it * pair_stride;
  if (tile >= static_cast<int64_t>(nmb) * nnb) return false;
  m_blk = static_cast<int>(tile / nnb);
  n_blk = static_cast<int>(tile - static_cast<int64_t>(m_blk) * nnb);
  return true;
}

template <int EPI, int EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1)
__launch_bounds__(PairCfg<EPI_WARPS, EPI == kEpiResidualLn>::kThreads, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out,
                         const GemmParams p) {
  using C = PairCfg<EPI_WARPS, EPI == kEpiResidualLn>;
  using ET = EpiTraits<EPI>;
  constexpr int kStages = C::kStages;
  constexpr int BLOCK_N = kPairBlockN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + kStages * C::kStageBytes;
  const uint32_t bar_base = staging_base + C::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };          // used in the leader CTA only
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };  // leader only
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  uint32_t* tmem_ptr_generic =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_ptr_smem - ptx::smem_u32(smem_raw)));

  constexpr int kExtraWarps = (EPI == kEpiResidualLn) ? kLnWarps : 0;  // LayerNorm warps sit after the epilogue warps
  constexpr int kTmaWarp = EPI_WARPS + kExtraWarps, kMmaWarp = kTmaWarp + 1;  // highest warp ids: never starved
  const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool is_leader = cta_rank == 0;

  if (warp_idx == kTmaWarp && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if constexpr (ET::kStaged) ptx::prefetch_tmap(&tmap_out);
    if constexpr (EPI == kEpiResidualLn) {
    }
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);   // leader's producer arms it with both CTAs' bytes
      ptx::mbar_init(empty_bar(s), 1);  // one multicast tcgen05.commit arrival
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar(a), 1);
      ptx::mbar_init(tmem_empty_bar(a), 2 * EPI_WARPS);  // one arrival per epilogue warp of BOTH CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == kMmaWarp) {
    ptx::tmem_alloc_pair<C::kTmemCols>(tmem_ptr_smem);
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // barriers of both CTAs initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  const int kseg_blocks = p.K / kBlockK;
  const int num_k_blocks = p.split3 == 1 ? 3 * kseg_blocks : (p.split3 == 2 ? 2 * kseg_blocks : kseg_blocks);
  const int64_t pair_idx = blockIdx.x >> 1;
  const int64_t pair_stride = gridDim.x >> 1;
  int m_blk = 0, n_blk = 0;

  if (warp_idx == kTmaWarp) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t it = 0; pair_tile(it, pair_idx, pair_stride, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk); ++it) {
        const int a_row = m_blk * (2 * kBlockM) + static_cast<int>(cta_rank) * kBlockM;
        const int b_row = n_blk * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / 2);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          int a_k, b_k;
          if (p.split3 == 1) {  // Ah*Wh, Ah*Wl, Al*Wh
            const int seg = kb / kseg_blocks;
            const int r = kb - seg * kseg_blocks;
            a_k = (seg == 2 ? p.K : 0) + r * kBlockK;
            b_k = (seg == 1 ? p.K : 0) + r * kBlockK;
          } else if (p.split3 == 2) {  // A is plain bf16: A*Wh, A*Wl
            const int seg = kb / kseg_blocks;
            const int r = kb - seg * kseg_blocks;
            a_k = r * kBlockK;
            b_k = seg * p.K + r * kBlockK;
          } else {
            a_k = b_k = kb * kBlockK;
          }
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
          const uint32_t leader_full = ptx::mapa(full_bar(stage), 0);
          if (is_leader) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * C::kStageBytes);
          ptx::tma_load_2d_pair(sa, &tmap_a, leader_full, a_k, a_row);
          ptx::tma_load_2d_pair(sb, &tmap_b, leader_full, b_k, b_row);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == kMmaWarp) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (is_leader && lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(2 * kBlockM, BLOCK_N) & p.idesc_mask;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t it = 0; pair_tile(it, pair_idx, pair_stride, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk); ++it) {
        ptx::mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
          const uint64_t desc_a = ptx::make_smem_desc_sw128(sa);
          const uint64_t desc_b = ptx::make_smem_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            ptx::umma_bf16_pair(tmem_d, desc_a + static_cast<uint64_t>(2 * k),
                                desc_b + static_cast<uint64_t>(2 * k), idesc,
                                (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit_pair(empty_bar(stage), 0b11);  // frees the slot in both CTAs
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit_pair(tmem_full_bar(acc), 0b11);  // accumulator halves ready in both CTAs
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (EPI == kEpiResidualLn && warp_idx >= EPI_WARPS) {
    // ===================== LayerNorm warps (fused LN only): warp j serves epilogue warp j =====================
    const int j = warp_idx - EPI_WARPS;
    const uint32_t want = static_cast<uint32_t>(p.num_n_blocks);
    for (int64_t it = 0; pair_tile(it, pair_idx, pair_stride, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk); ++it) {
      if (n_blk != p.num_n_blocks - 1) continue;  // the owner of the last N tile normalises the panel
      const int64_t mb = m_blk;
      if (lane == 0) {
        uint32_t* cnt = p.ln_sync + (mb * 8 + cta_rank * 4 + j);
        uint32_t have = 0, spins = 0;
        while (true) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(have) : "l"(cnt) : "memory");
          if (have >= want) break;
          __nanosleep(256);
          if (++spins > (1u << 23)) __trap();  // seconds without progress: fail instead of hanging
        }
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(cnt), "r"(0u) : "memory");  // ready for the next launch
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      __syncwarp();
      const int64_t row_begin = mb * (2 * kBlockM) + static_cast<int64_t>(cta_rank) * kBlockM + j * 32;
      const int64_t left = p.M - row_begin;
      const int nrows = left >= 32 ? 32 : static_cast<int>(left);
      if (nrows > 0) {
        switch (p.N) {
          case 384: ln_rows_from_l2<3>(p, row_begin, nrows, lane); break;
          case 768: ln_rows_from_l2<6>(p, row_begin, nrows, lane); break;
          default: ln_rows_from_l2<8>(p, row_begin, nrows, lane); break;  // 1024
        }
      }
    }
  } else {
    // ===================== epilogue warps (0..EPI_WARPS-1, both CTAs) =====================
    const int quarter = warp_idx & 3;             // TMEM lane quarter (rows) of this warp
    const int col_part = warp_idx >> 2;           // which slice of the 256 columns (EPI_WARPS == 8)
    constexpr int kColsPerWarp = BLOCK_N / (EPI_WARPS / 4);
    const uint32_t stg = staging_base + static_cast<uint32_t>(warp_idx) * (C::kStagingBufs * 32u * 128u);
    uint32_t stg_buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int ln_prev_m = -1;  // tile whose reduce-adds are issued but not yet known to be complete
    auto ln_publish = [&](int mb) {  // lane 0: this warp's slab of tile (mb, *) is final in L2
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.ln_sync + (static_cast<int64_t>(mb) * 8 + cta_rank * 4 + quarter)) : "memory");
    };
    (void)ln_prev_m;
    for (int64_t it = 0; pair_tile(it, pair_idx, pair_stride, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk); ++it) {
      const int row0 = m_blk * (2 * kBlockM) + static_cast<int>(cta_rank) * kBlockM + quarter * 32;
      const int n0 = n_blk * BLOCK_N;
      ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
      const uint32_t leader_empty = ptx::mapa(tmem_empty_bar(acc), 0);
      constexpr int kTileEpi = (EPI == kEpiResidualLn) ? kEpiResidualTma : EPI;
      epilogue_tile<kTileEpi, C::kStagingBufs>(p, &tmap_out, taddr, row0, lane, n0, col_part * kColsPerWarp,
                                               (col_part + 1) * kColsPerWarp, stg, stg_buf,
                         [&]() {
                           __syncwarp();
                           if (lane == 0) ptx::mbar_arrive_cluster(leader_empty);
                         });
      if constexpr (EPI == kEpiResidualLn) {
        if (ln_prev_m >= 0 && lane == 0) {
          // everything older than this tile's BLOCK_N / 32 reduce-add groups has completed
          ptx::tma_store_wait<BLOCK_N / 32>();
          ln_publish(ln_prev_m);
        }
        ln_prev_m = m_blk;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if constexpr (ET::kStaged) {
      if (lane == 0) ptx::tma_store_wait<0>();
    }
    if constexpr (EPI == kEpiResidualLn) {
      if (ln_prev_m >= 0 && lane == 0) ln_publish(ln_prev_m);
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();  // peer may still be reading our smem / signalling our barriers
  if (warp_idx == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair<C::kTmemCols>(tmem_base);
  }
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// Row-major [rows, cols] matrix (leading dim ld elements); box = box_rows x (128 B of columns),
// 128B swizzle.  elem_bytes 2 -> bf16, 4 -> fp32.
int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld,
              int box_rows, int elem_bytes) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DUO_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
    return DUO_ERR_CUDA;
  }
  return DUO_OK;
}

template <int BLOCK_N, int EPI>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmParams& p,
           cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  static uint64_t configured = 0;  // per device
  auto kfn = gemm_tcgen05_kernel<BLOCK_N, EPI>;
  if (first_use_on_device(configured))
    DUO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(C::kSmemBytes)));
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;
  const int sms = device_sm_count();
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  kfn<<<grid, kNumThreads, C::kSmemBytes, st>>>(ta, tb, to, p);
  DUO_LAUNCH_CHECK("gemm_tcgen05_kernel");
  return DUO_OK;
}

template <int BLOCK_N>
int dispatch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                 const GemmParams& p, int epi, cudaStream_t st) {
  switch (epi) {
    case DUO_EPI_BF16: return launch<BLOCK_N, DUO_EPI_BF16>(ta, tb, to, p, st);
    case DUO_EPI_GELU_BF16: return launch<BLOCK_N, DUO_EPI_GELU_BF16>(ta, tb, to, p, st);
    case DUO_EPI_RESIDUAL_F32: return launch<BLOCK_N, DUO_EPI_RESIDUAL_F32>(ta, tb, to, p, st);
    case kEpiResidualTma: return launch<BLOCK_N, kEpiResidualTma>(ta, tb, to, p, st);
    case DUO_EPI_SCATTER_F32: return launch<BLOCK_N, DUO_EPI_SCATTER_F32>(ta, tb, to, p, st);
    case DUO_EPI_F32: return launch<BLOCK_N, DUO_EPI_F32>(ta, tb, to, p, st);
    case DUO_EPI_SPLIT_BF16: return launch<BLOCK_N, DUO_EPI_SPLIT_BF16>(ta, tb, to, p, st);
    case DUO_EPI_GELU_SPLIT_BF16: return launch<BLOCK_N, DUO_EPI_GELU_SPLIT_BF16>(ta, tb, to, p, st);
    default: set_error("duo_gemm: unknown epilogue %d", epi); return DUO_ERR_INVALID;
  }
}

template <int EPI, int EPI_WARPS>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                const GemmParams& p, cudaStream_t st) {
  using C = PairCfg<EPI_WARPS, EPI == kEpiResidualLn>;
  static uint64_t configured = 0;  // per device
  auto kfn = gemm_tcgen05_pair_kernel<EPI, EPI_WARPS>;
  if (first_use_on_device(configured))
    DUO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(C::kSmemBytes)));
  const int64_t units = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;  // tiles, round-robin over pairs
  // DUO_GEMM_MAX_SMS caps the persistent grid (leaves SMs to kernels running beside the GEMM)
  static const int sm_cap = [] { const char* e = getenv("DUO_GEMM_MAX_SMS"); return e ? atoi(e) : 0; }();
  const int sms_avail = device_sm_count();
  const int pairs_max = ((sm_cap > 1 && sm_cap < sms_avail) ? sm_cap : sms_avail) / 2;
  const int pairs = static_cast<int>(units < pairs_max ? units : pairs_max);
  kfn<<<2 * pairs, C::kThreads, C::kSmemBytes, st>>>(ta, tb, to, p);
  DUO_LAUNCH_CHECK("gemm_tcgen05_pair_kernel");
  return DUO_OK;
}

int dispatch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                  const GemmParams& p, int epi, cudaStream_t st) {
  switch (epi) {
    case kEpiResidualLn: return launch_pair<kEpiResidualLn, 4>(ta, tb, to, p, st);
    case DUO_EPI_BF16: return launch_pair<DUO_EPI_BF16, 4>(ta, tb, to, p, st);
    case DUO_EPI_GELU_BF16: {
      static const int gelu_warps = [] { const char* e = getenv("DUO_GEMM_GELU_WARPS"); return (e && e[0] == '4') ? 4 : 8; }();
      return gelu_warps == 8 ? launch_pair<DUO_EPI_GELU_BF16, 8>(ta, tb, to, p, st)
                             : launch_pair<DUO_EPI_GELU_BF16, 4>(ta, tb, to, p, st);
    }
    case DUO_EPI_RESIDUAL_F32: return launch_pair<DUO_EPI_RESIDUAL_F32, 4>(ta, tb, to, p, st);
    case kEpiResidualTma: return launch_pair<kEpiResidualTma, 4>(ta, tb, to, p, st);
    case DUO_EPI_SCATTER_F32: {  // short K, store-bound epilogue: 8 warps (DUO_GEMM_SCATTER_WARPS=4 for the A/B)
      static const int w = [] { const char* e = getenv("DUO_GEMM_SCATTER_WARPS"); return (e && e[0] == '4') ? 4 : 8; }();
      return w == 8 ? launch_pair<DUO_EPI_SCATTER_F32, 8>(ta, tb, to, p, st)
                    : launch_pair<DUO_EPI_SCATTER_F32, 4>(ta, tb, to, p, st);
    }
    case DUO_EPI_F32: return launch_pair<DUO_EPI_F32, 4>(ta, tb, to, p, st);
    case DUO_EPI_SPLIT_BF16: return launch_pair<DUO_EPI_SPLIT_BF16, 4>(ta, tb, to, p, st);
    case DUO_EPI_GELU_SPLIT_BF16: return launch_pair<DUO_EPI_GELU_SPLIT_BF16, 8>(ta, tb, to, p, st);
    default: set_error("duo_gemm: unknown epilogue %d", epi); return DUO_ERR_INVALID;
  }
}

// DUO_GEMM_PAIR=0 disables the CTA-pair (cta_group::2) kernel.
bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DUO_GEMM_PAIR");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// DUO_GEMM_RESIDUAL=direct selects the load/add/store residual epilogue instead of TMA reduce-add.
bool residual_via_tma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DUO_GEMM_RESIDUAL");
    v = (e != nullptr && e[0] == 'd') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace
}  // namespace duo

extern "C" int duo_gemm(const duo_gemm_args* a, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(a != nullptr, "duo_gemm: args is NULL");
  DUO_CHECK_ARG(a->A && a->W && a->out, "duo_gemm: NULL operand");
  DUO_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "duo_gemm: empty problem M=%lld N=%d K=%d",
                (long long)a->M, a->N, a->K);
  DUO_CHECK_ARG(a->N % 128 == 0, "duo_gemm: N=%d must be a multiple of 128", a->N);
  DUO_CHECK_ARG(a->K % kBlockK == 0, "duo_gemm: K=%d must be a multiple of 64", a->K);
  DUO_CHECK_ARG(a->split3 >= 0 && a->split3 <= 2, "duo_gemm: split3=%d", a->split3);
  DUO_CHECK_ARG(a->fp16_operands == 0 || (a->fp16_operands == 1 && a->split3 == 0), "duo_gemm: fp16_operands=%d needs split3 == 0", a->fp16_operands);
  const int kcols = a->split3 ? 2 * a->K : a->K;         // W columns
  const int acols = a->split3 == 1 ? 2 * a->K : a->K;    // A columns
  DUO_CHECK_ARG(a->lda >= acols && a->ldw >= kcols && a->lda % 8 == 0 && a->ldw % 8 == 0,
                "duo_gemm: bad leading dims lda=%lld ldw=%lld", (long long)a->lda, (long long)a->ldw);
  DUO_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                "duo_gemm: operands must be 16-byte aligned");
  DUO_CHECK_ARG(a->ldo % 8 == 0, "duo_gemm: ldo=%lld must be a multiple of 8", (long long)a->ldo);
  if (a->epilogue == DUO_EPI_SCATTER_F32) {
    DUO_CHECK_ARG(a->row_map && a->rows_per_group > 0 && a->dest_rows_per_group > 0,
                  "duo_gemm: scatter epilogue needs row_map / group sizes");
    DUO_CHECK_ARG(a->pos == nullptr || a->pos_period > 0, "duo_gemm: pos_period must be > 0");
  }
  DUO_CHECK_ARG(a->M < (int64_t(1) << 31) - 256, "duo_gemm: M too large for a 32-bit TMA coordinate");

  // Tile shape: CTA-pair 256x256 when N allows and there are at least two waves of pair tiles;
  // else 128 x 256 (enough tiles to fill the machine) or 128 x 128.
  int64_t m_blocks = (a->M + kBlockM - 1) / kBlockM;
  int block_n = 128;
  if (a->N % 256 == 0 && m_blocks * (a->N / 256) >= 2 * device_sm_count()) block_n = 256;
  const int64_t m2_blocks = (a->M + 2 * kBlockM - 1) / (2 * kBlockM);
  const bool use_pair = pair_enabled() && a->N % 256 == 0 && m2_blocks * (a->N / 256) >= device_sm_count();
  if (use_pair) {
    block_n = 128;  // W box rows: each CTA loads half of the 256-wide tile
    m_blocks = m2_blocks;
  }

  CUtensorMap ta, tb, to;
  int rc = make_tmap(&ta, a->A, a->M, acols, a->lda, kBlockM, 2);
  if (rc != DUO_OK) return rc;
  rc = make_tmap(&tb, a->W, a->N, kcols, a->ldw, block_n, 2);
  if (rc != DUO_OK) return rc;
  int epi = a->epilogue;
  const bool want_ln = a->ln_out != nullptr;
  if (want_ln) {
    DUO_CHECK_ARG(epi == DUO_EPI_RESIDUAL_F32 && !a->split3 && a->ln_gamma && a->ln_beta,
                  "duo_gemm: fused LayerNorm needs the bf16 residual epilogue and ln_gamma / ln_beta");
    DUO_CHECK_ARG(a->N % 128 == 0 && a->N <= 1024, "duo_gemm: fused LayerNorm needs N %% 128 == 0, N <= 1024");
    DUO_CHECK_ARG((reinterpret_cast<uintptr_t>(a->ln_out) & 15) == 0, "duo_gemm: ln_out must be 16-byte aligned");
  }
  const bool fused_ln = want_ln && use_pair && a->ln_sync != nullptr && (a->N == 384 || a->N == 768 || a->N == 1024);
  if (fused_ln) epi = kEpiResidualLn;
  if (epi == DUO_EPI_RESIDUAL_F32 && residual_via_tma()) epi = kEpiResidualTma;
  if (epi == DUO_EPI_BF16 || epi == DUO_EPI_GELU_BF16) {
    rc = make_tmap(&to, a->out, a->M, a->N, a->ldo, 32, 2);
  } else if (epi == kEpiResidualTma || epi == kEpiResidualLn) {
    rc = make_tmap(&to, a->out, a->M, a->N, a->ldo, 32, 4);
  } else {
    to = ta;  // unused by the direct-store epilogues
  }
  if (rc != DUO_OK) return rc;
  GemmParams p;
  p.ln_gamma = a->ln_gamma;
  p.ln_beta = a->ln_beta;
  p.ln_out = a->ln_out;
  p.ln_sync = a->ln_sync;
  p.ln_eps = a->ln_eps;
  p.relu = a->relu;
  p.bias = a->bias;
  p.out = a->out;
  p.gamma = a->gamma;
  p.row_map = a->row_map;
  p.pos = a->pos;
  p.M = a->M;
  p.ldo = a->ldo;
  p.N = a->N;
  p.K = a->K;
  p.split3 = a->split3;
  p.rows_per_group = a->rows_per_group;
  p.dest_rows_per_group = a->dest_rows_per_group;
  p.pos_period = a->pos_period;
  p.num_m_blocks = static_cast<int32_t>(m_blocks);
  p.idesc_mask = a->fp16_operands ? ~((1u << 7) | (1u << 10)) : ~0u;  // a_format / b_format: 1 = BF16, 0 = F16
  p.num_n_blocks = use_pair ? a->N / kPairBlockN : a->N / block_n;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (use_pair) return dispatch_pair(ta, tb, to, p, epi, st);
  rc = block_n == 256 ? dispatch_epi<256>(ta, tb, to, p, epi, st) : dispatch_epi<128>(ta, tb, to, p, epi, st);
  if (rc != DUO_OK || !want_ln) return rc;
  // small problems: unfused — LayerNorm of the updated rows as a second launch
  return duo_layernorm(reinterpret_cast<const float*>(a->out), a->ln_gamma, a->ln_beta, a->ln_out, DUO_ACT_BF16,
                       a->M, a->N, a->ldo, a->ln_eps, stream);
}
