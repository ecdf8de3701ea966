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
//               bias / GELU(erf) / LayerScale / a forwarded LayerNorm (rstd of the row, merged from the
//               statistics the producing GEMM left; W is then the row-centred W * diag(ln_weight)), then
//                 * bf16 outputs: rows staged in a per-warp 128B-swizzled shared-memory tile and
//                   written with TMA stores (full-line, coalesced);
//                 * residual (X += ...): staged fp32 tile + TMA reduce-add into the fp32 residual
//                   stream (the read-modify-write happens in L2, no SM-side loads);
//                 * residual with statistics forwarding (pair kernel): X chunks TMA-loaded into a per-warp
//                   ring, updated in place, TMA-stored together with bf16(x - previous row mean) and the
//                   row's (mean, M2) over the tile's 256 columns — the LayerNorm that follows needs no pass
//                   over X of its own;
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

// Internal epilogue variants (superset of the ABI's DUO_EPI_*).
constexpr int kEpiResidualTma = 100;  // DUO_EPI_RESIDUAL_F32: TMA reduce-add into the fp32 stream
constexpr int kEpiResidualFwd = 101;  // residual update + statistics forwarding (pair kernel): X is TMA-loaded,
                                      // updated in shared memory and stored back together with its bf16 copy and
                                      // per-row LayerNorm partial statistics (duo_gemm_args.xb_out / stats_out)
constexpr int kEpiResidualFwdLongK = 104;  // same, tuned for long K (fc2): one more operand stage, shallower X ring
constexpr int kEpiBf16Ln = 102;       // DUO_EPI_BF16 with the forwarded LayerNorm applied in the epilogue
constexpr int kEpiGeluBf16Ln = 103;   // DUO_EPI_GELU_BF16 with the forwarded LayerNorm applied in the epilogue

constexpr int kStatCols = 256;  // columns covered by one forwarded (mean, M2) pair (= the N tile of the producer)

template <int EPI>
struct EpiTraits {
  static constexpr bool kLnApply = (EPI == kEpiBf16Ln || EPI == kEpiGeluBf16Ln);
  static constexpr bool kGelu = (EPI == DUO_EPI_GELU_BF16 || EPI == kEpiGeluBf16Ln);
  static constexpr bool kStagedBf16 = (EPI == DUO_EPI_BF16 || EPI == DUO_EPI_GELU_BF16 || kLnApply);
  static constexpr bool kStagedF32 = (EPI == kEpiResidualTma);
  static constexpr bool kFwd = (EPI == kEpiResidualFwd || EPI == kEpiResidualFwdLongK);
  // X chunks in flight per epilogue warp (ring of in-place slots).  Short K (proj): the tile's MMAs take ~6 us and
  // its 128 KB of X per CTA must stream in meanwhile -> 3 chunks ahead per warp.  Long K (fc2): ~20 us per tile,
  // one chunk ahead is plenty and the shared memory buys a fifth operand stage instead.
  static constexpr int kFwdSlots = (EPI == kEpiResidualFwd) ? 4 : (EPI == kEpiResidualFwdLongK ? 2 : 0);
  static constexpr int kFwdStages = (EPI == kEpiResidualFwd) ? 4 : 5;
  static constexpr bool kStaged = kStagedBf16 || kStagedF32 || kFwd;
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
  // statistics forwarding, producer side (kEpiResidualFwd)
  float2* stats_out;        // [M, N / 256] (mean, M2) of the updated rows, per 256-column part
  const float2* shift_stats;  // [M, N / 256] statistics of the rows BEFORE this update (or NULL): the bf16 copy is x - mean_old
  // statistics forwarding, consumer side (kEpiBf16Ln / kEpiGeluBf16Ln)
  const float2* ln_stats;   // [M, K / 256]
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

// Forwarded LayerNorm, consumer side.  The A operand of this GEMM is the UN-normalised bf16 copy of the
// residual stream and W'' = W * diag(ln_gamma) with every ROW CENTRED (sum_k W''[n, k] = 0, engine.pack_ln_linear):
// the mean of x then cancels inside the tensor-core product, x W''^T = (x - mean) (W diag(gamma))^T, and
//   LN(x) W^T + b = rstd * (x W''^T) + (W ln_beta + b)
// so the epilogue computes  ln_a * acc + bias'[n]  with ln_a = rstd — one FMA where the plain epilogue has an add.
// The (mean, M2) pairs of the row's K / 256 column parts (written by the producing residual GEMM from the
// fp32 row) are merged with Chan's formula.
constexpr int kMaxStatParts = 4;  // K <= 1024
struct LnRowStats {
  float2 s[kMaxStatParts];
};
// Issue the loads of one row's partial statistics (done one tile ahead: the latency of these L2 / DRAM reads
// hides behind the epilogue of the current tile).
__device__ __forceinline__ void ln_stats_load(const GemmParams& p, int64_t row, LnRowStats& st) {
  const int parts = p.K / kStatCols;
  const bool valid = row < p.M;
  const float2* src = p.ln_stats + row * parts;
#pragma unroll
  for (int i = 0; i < kMaxStatParts; ++i)
    st.s[i] = (valid && i < parts) ? __ldg(src + i) : make_float2(0.f, 0.f);
}
__device__ __forceinline__ void ln_stats_finish(const GemmParams& p, const LnRowStats& st, float& ln_a) {
  const int parts = p.K / kStatCols;
  float mean = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxStatParts; ++i) mean += st.s[i].x;  // absent parts are zero
  mean *= 1.0f / static_cast<float>(parts);
#pragma unroll
  for (int i = 0; i < kMaxStatParts; ++i) {
    if (i < parts) {
      const float d = st.s[i].x - mean;
      m2 += st.s[i].y + static_cast<float>(kStatCols) * d * d;
    }
  }
  ln_a = rsqrtf(m2 / static_cast<float>(p.K) + p.ln_eps);
}

// Bias slice [col, col + 4 * NQ) -> registers; issued BEFORE waiting on the TMEM load so both latencies overlap.
template <int NQ>
__device__ __forceinline__ void epilogue_bias_load(const GemmParams& p, int col, float4 (&b)[NQ]) {
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
    for (int j = 0; j < NQ; ++j) b[j] = __ldg(b4 + j);
  } else {
#pragma unroll
    for (int j = 0; j < NQ; ++j) b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// f = ln_a * acc + b  (ln_a == 1 without a forwarded LayerNorm: plain bias add), then the epilogue's function;
// 4 * NQ consecutive columns starting at `col`
template <int EPI, int NQ>
__device__ __forceinline__ void epilogue_math(const GemmParams& p, int col, const uint32_t (&v)[4 * NQ],
                                              const float4 (&b)[NQ], float (&f)[4 * NQ], float ln_a) {
  using ET = EpiTraits<EPI>;
  if constexpr (ET::kGelu) {  // bias add and GELU on packed fp32 pairs
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      uint64_t lo, hi;
      if constexpr (ET::kLnApply) {
        lo = fma2(pack2(ln_a, ln_a), pack2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack2(b[j].x, b[j].y));
        hi = fma2(pack2(ln_a, ln_a), pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack2(b[j].z, b[j].w));
      } else {
        lo = add2(pack2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack2(b[j].x, b[j].y));
        hi = add2(pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack2(b[j].z, b[j].w));
      }
#ifndef DUO_GELU_TANH_FORM
#define DUO_GELU_TANH_FORM 1  // one-MUFU tanh form; 0 = sigmoid form (ex2 + rcp), csrc/Makefile `tuning`: same speed alone
#endif                        // (0.931 vs 0.933 ms per 64 images), 0 - 1.5 % slower inside the step (profiles/r02_notes.md)
#if DUO_GELU_TANH_FORM
      unpack2(gelu_erf_tanh_p2(lo), f[4 * j + 0], f[4 * j + 1]);
      unpack2(gelu_erf_tanh_p2(hi), f[4 * j + 2], f[4 * j + 3]);
#else
      unpack2(gelu_erf_sigmoid_p2(lo), f[4 * j + 0], f[4 * j + 1]);
      unpack2(gelu_erf_sigmoid_p2(hi), f[4 * j + 2], f[4 * j + 3]);
#endif
    }
    return;
  }
  // (packed fp32 pairs: half the issue slots of the scalar form)
  if constexpr (ET::kLnApply) {
    const uint64_t a2 = pack2(ln_a, ln_a);
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      unpack2(fma2(a2, pack2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack2(b[j].x, b[j].y)),
              f[4 * j + 0], f[4 * j + 1]);
      unpack2(fma2(a2, pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack2(b[j].z, b[j].w)),
              f[4 * j + 2], f[4 * j + 3]);
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    unpack2(add2(pack2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack2(b[j].x, b[j].y)),
            f[4 * j + 0], f[4 * j + 1]);
    unpack2(add2(pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack2(b[j].z, b[j].w)),
            f[4 * j + 2], f[4 * j + 3]);
  }
  if constexpr (EPI == DUO_EPI_BF16 || EPI == DUO_EPI_F32) {
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 4 * NQ; ++j) f[j] = fmaxf(f[j], 0.f);
    }
  }
  if constexpr (EPI == DUO_EPI_GELU_SPLIT_BF16) {
#pragma unroll
    for (int j = 0; j < 4 * NQ; ++j) f[j] = gelu_erf(f[j]);
  }
  if constexpr (EPI == kEpiResidualTma) {
    if (p.gamma != nullptr) {
      const float4* g4 = reinterpret_cast<const float4*>(p.gamma + col);
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
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
template <int EPI, int NBUF, int SUBCOLS = 32, typename ReleaseFn>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const CUtensorMap* tmap_out,
                                              uint32_t taddr, int row0, int lane, int n0, int c_begin,
                                              int c_end, uint32_t stg, uint32_t& stg_buf,
                                              float ln_a, ReleaseFn release) {
  using ET = EpiTraits<EPI>;
  const int64_t row = static_cast<int64_t>(row0) + lane;
  const bool valid = row < p.M;
  const uint32_t my_row_off = static_cast<uint32_t>(lane) * 128u;
  (void)valid;
  (void)my_row_off;
  if constexpr (ET::kStagedBf16 && SUBCOLS == 16) {
    // 16-warp epilogue (GELU): each warp owns 64 columns = ONE 128-byte-wide staging tile per accumulator, filled 16
    // columns at a time (tcgen05.ld x16: half the live registers, so four epilogue warps fit on every scheduler);
    // single staging buffer: the TMA store of the previous tile was issued a whole tile ago.
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 64) {
      const uint32_t buf = stg + my_row_off;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint32_t v[16];
        float4 bia[4];
        ptx::tmem_ld_32x16(taddr + static_cast<uint32_t>(c + 16 * h), v);
        epilogue_bias_load(p, n0 + c + 16 * h, bia);
        ptx::tmem_ld_wait();
        if (h == 3 && c + 64 >= c_end) {  // accumulator fully read: hand the TMEM buffer back early
          ptx::tc_fence_before();
          release();
        }
        float f[16];
        epilogue_math<EPI>(p, n0 + c + 16 * h, v, bia, f, ln_a);
        if (h == 0) {
          if (lane == 0) ptx::tma_store_wait_read<0>();  // the staging tile is no longer being read
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)  // 16-byte chunk (2h + j) of this row, XOR-swizzled
          st_shared_v4(buf + (static_cast<uint32_t>((2 * h + j) ^ (lane & 7)) << 4),
                       pack_bf16x2(f[8 * j + 0], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                       pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(tmap_out, stg, n0 + c, row0);
        ptx::tma_store_commit();
      }
    }
  } else if constexpr (ET::kStagedBf16) {
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
        epilogue_math<EPI>(p, n0 + c + 32 * h, v, bia, f, ln_a);
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
        // (an L2 evict-first hint on these stores: no difference, ABAB 194.4 vs 194.6 ms per step)
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
      epilogue_math<EPI>(p, n0 + c, v, bia, f, 1.f);
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
      epilogue_math<EPI>(p, n0 + c, v, bia, f, 1.f);
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
        epilogue_math<EPI>(p, n0 + c, v, bia, f, 1.f);
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
    LnRowStats ln_next;  // forwarded LayerNorm: the row statistics of the NEXT tile, loaded one tile ahead
    if constexpr (ET::kLnApply) {
      if (static_cast<int64_t>(blockIdx.x) < num_tiles)
        ln_stats_load(p, static_cast<int64_t>(blockIdx.x / p.num_n_blocks) * kBlockM + quarter * 32 + lane, ln_next);
    }
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = static_cast<int>(tile / p.num_n_blocks);
      const int n_blk = static_cast<int>(tile - static_cast<int64_t>(m_blk) * p.num_n_blocks);
      const int row0 = m_blk * kBlockM + quarter * 32;  // first row of this warp's slab
      const int n0 = n_blk * BLOCK_N;
      float ln_a = 1.f;
      if constexpr (ET::kLnApply) {
        ln_stats_finish(p, ln_next, ln_a);
        const int64_t nt = tile + gridDim.x;
        if (nt < num_tiles) ln_stats_load(p, (nt / p.num_n_blocks) * kBlockM + quarter * 32 + lane, ln_next);
      }
      ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);

      epilogue_tile<EPI, 2>(p, &tmap_out, taddr, row0, lane, n0, 0, BLOCK_N, stg, stg_buf, ln_a,
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
// Statistics-forwarding residual epilogue (kEpiResidualFwd*): per epilogue warp a ring of kFwdSlots
// slots, each one 32 x 32 fp32 chunk of X (TMA-loaded, updated in place, TMA-stored) plus its bf16 copy.
constexpr uint32_t kFwdXBytes = 32 * 128;  // 32 rows x 32 fp32, SWIZZLE_128B
constexpr uint32_t kFwdBBytes = 32 * 64;   // 32 rows x 32 bf16, SWIZZLE_64B
// EPI_WARPS epilogue warps: 4 (one per TMEM lane quarter, all 256 columns, 6 operand stages), 8 (two per
// quarter, 128 columns each, 5 operand stages: GELU, token scatter) or 16 (four per quarter, 64 columns each in
// 16-column steps, < 96 registers per thread; a tuning variant of the GELU epilogue: fc1 + GELU 0.967 ms per 64
// images against 0.950 ms with 8 warps — the epilogue is bound by the MUFU pipe (ncu: XU 60 % busy, stalls `wait` /
// `mio_throttle`), not by the number of warps in flight).
template <int EPI, int EPI_WARPS>
struct PairCfg {
  using ET = EpiTraits<EPI>;
  static constexpr bool FWD = ET::kFwd;
  static constexpr int kFwdSlots = ET::kFwdSlots;
  // (8 warps with ONE staging tile each and a sixth operand stage: no difference for fc1 + GELU, ABAB 194.4 vs 196.3 ms per step)
  static constexpr bool kSingleStaging = EPI_WARPS == 16;
  static constexpr int kStages = FWD ? ET::kFwdStages : (EPI_WARPS >= 8 ? 5 : 6);
  static constexpr int kThreads = 64 + 32 * EPI_WARPS;
  static constexpr int kStagingBufs = kSingleStaging ? 1 : 2;  // 64 KB of staging for 8 and for 16 warps
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;          // 16 KB (this CTA's 128 rows)
  static constexpr uint32_t kBBytes = (kPairBlockN / 2) * kBlockK * 2;  // 16 KB (this CTA's N half)
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * kPairBlockN;
  static constexpr uint32_t kStagingBytes =
      FWD ? EPI_WARPS * kFwdSlots * (kFwdXBytes + kFwdBBytes) : EPI_WARPS * kStagingBufs * 32 * 128;
  static constexpr int kNumBars = 2 * kStages + 4;                    // + one slot for the TMEM pointer
  static constexpr int kNumXBars = FWD ? EPI_WARPS * kFwdSlots : 0;   // "X chunk loaded", one per ring slot
  static constexpr uint32_t kBarrierBytes = (kNumBars + 1 + kNumXBars) * 8;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory budget of one CTA per SM exceeded");
};

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

__device__ __forceinline__ void ld_shared_v4(uint32_t addr, float& a, float& b, float& c, float& d) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr) : "memory");
}

template <int EPI, int EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1)
__launch_bounds__(PairCfg<EPI, EPI_WARPS>::kThreads, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out,
                         const __grid_constant__ CUtensorMap tmap_aux,  // kEpiResidualFwd: the bf16 copy
                         const GemmParams p) {
  using C = PairCfg<EPI, EPI_WARPS>;
  using ET = EpiTraits<EPI>;
  constexpr int kStages = C::kStages;
  constexpr int BLOCK_N = kPairBlockN;
  constexpr int kFwdSlots = C::kFwdSlots;
  (void)kFwdSlots;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + kStages * C::kStageBytes;
  const uint32_t bar_base = staging_base + C::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };          // used in the leader CTA only
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };  // leader only
  const uint32_t tmem_ptr_smem = bar_base + 8u * C::kNumBars;
  auto x_bar = [&](int i) { return bar_base + 8u * (C::kNumBars + 1 + i); };  // kEpiResidualFwd only
  (void)x_bar;
  uint32_t* tmem_ptr_generic =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_ptr_smem - ptx::smem_u32(smem_raw)));

  constexpr int kTmaWarp = EPI_WARPS, kMmaWarp = kTmaWarp + 1;  // highest warp ids: never starved
  const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool is_leader = cta_rank == 0;

  if (warp_idx == kTmaWarp && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if constexpr (ET::kStaged) ptx::prefetch_tmap(&tmap_out);
    if constexpr (ET::kFwd) ptx::prefetch_tmap(&tmap_aux);
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
    for (int i = 0; i < C::kNumXBars; ++i) ptx::mbar_init(x_bar(i), 1);
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
  } else {
    // ===================== epilogue warps (0..EPI_WARPS-1, both CTAs) =====================
    const int quarter = warp_idx & 3;             // TMEM lane quarter (rows) of this warp
    const int col_part = warp_idx >> 2;           // which slice of the 256 columns (EPI_WARPS == 8)
    constexpr int kColsPerWarp = BLOCK_N / (EPI_WARPS / 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (ET::kFwd) {
      // ---- residual update with statistics forwarding --------------------------------------------
      // X chunk g of this warp (32 rows x 32 columns, chunks numbered across tiles) lives in ring slot
      // g % kFwdSlots: TMA load (issued kFwdSlots - 1 chunks ahead, also across tile boundaries, so the
      // reads of X overlap the MMAs of the tile) -> x += gamma * (acc + bias) in place -> TMA store of the
      // fp32 chunk and of its bf16 copy (one bulk group).  A slot is reloaded once the group that stored
      // it has been read out (wait_group.read 1 right after committing the NEXT group).
      constexpr int kChunks = kColsPerWarp / 32;
      static_assert(kColsPerWarp == kStatCols, "one (mean, M2) pair per epilogue warp and tile");
      const uint32_t xbuf0 = staging_base + static_cast<uint32_t>(warp_idx) * (kFwdSlots * kFwdXBytes);
      const uint32_t bbuf0 = staging_base + EPI_WARPS * kFwdSlots * kFwdXBytes +
                             static_cast<uint32_t>(warp_idx) * (kFwdSlots * kFwdBBytes);
      const int parts = p.N / kStatCols;
      auto issue_load = [&](uint32_t gi) {  // lane 0 only
        int mb, nb;
        if (!pair_tile(gi / kChunks, pair_idx, pair_stride, p.num_m_blocks, p.num_n_blocks, mb, nb)) return;
        const int r = mb * (2 * kBlockM) + static_cast<int>(cta_rank) * kBlockM + quarter * 32;
        const int c = nb * BLOCK_N + col_part * kColsPerWarp + static_cast<int>(gi % kChunks) * 32;
        const uint32_t s = gi % kFwdSlots;
        const uint32_t bar = x_bar(warp_idx * kFwdSlots + static_cast<int>(s));
        ptx::mbar_arrive_expect_tx(bar, kFwdXBytes);
        // (an L2 evict-first hint on these read-once loads was measured: proj 0.424 -> 0.481 ms per 64 images, fc2 unchanged)
        ptx::tma_load_2d(xbuf0 + s * kFwdXBytes, &tmap_out, bar, c, r);
      };
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i + 1 < kFwdSlots; ++i) issue_load(static_cast<uint32_t>(i));
      }
      uint32_t g = 0;
      const uint32_t sw128 = static_cast<uint32_t>(lane & 7);         // 16-byte chunk XOR, 128 B rows
      const uint32_t sw64 = static_cast<uint32_t>((lane >> 1) & 3);   // 16-byte chunk XOR, 64 B rows
      for (int64_t it = 0; pair_tile(it, pair_idx, pair_stride, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk); ++it) {
        const int row0 = m_blk * (2 * kBlockM) + static_cast<int>(cta_rank) * kBlockM + quarter * 32;
        const int n0 = n_blk * BLOCK_N;
        const int64_t row = static_cast<int64_t>(row0) + lane;
        const bool valid = row < p.M;
        // Row shift of the bf16 copy: the row mean as of the PREVIOUS statistics.  The consumer's weights are
        // row-centred, so any per-row constant drops out of its product; subtracting (an estimate of) the mean before
        // rounding keeps the bf16 rounding error relative to the row's spread instead of its offset.
        float shift = 0.f;
        if (p.shift_stats != nullptr && valid) {
          for (int i = 0; i < parts; ++i) shift += __ldg(p.shift_stats + row * parts + i).x;
          shift *= 1.0f / static_cast<float>(parts);
        }
        ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
        const uint32_t leader_empty = ptx::mapa(tmem_empty_bar(acc), 0);
        uint64_t s1 = 0, s2 = 0;  // packed (even, odd) partial sums of (x - shift), (x - shift)^2
#pragma unroll 1
        for (int ci = 0; ci < kChunks; ++ci, ++g) {
          const int c = col_part * kColsPerWarp + ci * 32;  // first column of the chunk inside the tile
          const uint32_t s = g % kFwdSlots;
          const uint32_t xrow = xbuf0 + s * kFwdXBytes + static_cast<uint32_t>(lane) * 128u;
          const uint32_t brow = bbuf0 + s * kFwdBBytes + static_cast<uint32_t>(lane) * 64u;
          uint32_t v[32];
          float4 bia[8];
          ptx::tmem_ld_32x32(taddr + static_cast<uint32_t>(c), v);
          epilogue_bias_load(p, n0 + c, bia);
          ptx::mbar_wait(x_bar(warp_idx * kFwdSlots + static_cast<int>(s)), (g / kFwdSlots) & 1u);
          ptx::tmem_ld_wait();
          if (ci == kChunks - 1) {  // accumulator fully read: hand the TMEM buffer back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(leader_empty);
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x0, x1, x2, x3;
            ld_shared_v4(xrow + ((static_cast<uint32_t>(j) ^ sw128) << 4), x0, x1, x2, x3);
            const uint64_t alo = add2(pack2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack2(bia[j].x, bia[j].y));
            const uint64_t ahi = add2(pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack2(bia[j].z, bia[j].w));
            if (p.gamma != nullptr) {
              const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + n0 + c) + j);
              unpack2(fma2(pack2(gm.x, gm.y), alo, pack2(x0, x1)), f[4 * j + 0], f[4 * j + 1]);
              unpack2(fma2(pack2(gm.z, gm.w), ahi, pack2(x2, x3)), f[4 * j + 2], f[4 * j + 3]);
            } else {
              unpack2(add2(pack2(x0, x1), alo), f[4 * j + 0], f[4 * j + 1]);
              unpack2(add2(pack2(x2, x3), ahi), f[4 * j + 2], f[4 * j + 3]);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)  // updated fp32 chunk, in place
            st_shared_v4(xrow + ((static_cast<uint32_t>(j) ^ sw128) << 4), __float_as_uint(f[4 * j + 0]),
                         __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
          // d = x - shift feeds both the bf16 copy (A operand of the next GEMM) and the LayerNorm partial statistics
          // of the row over the tile's 256 columns: sums of d and d^2 — the previous row mean is the ideal pivot of
          // the shifted-data variance (no cancellation for |mean| >> spread); without shift_stats these are raw sums
          if (ci == 0) {
            s1 = 0;
            s2 = 0;
          }
          {
            const uint64_t ns = pack2(-shift, -shift);
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint64_t d = add2(pack2(f[2 * j], f[2 * j + 1]), ns);
              s1 = add2(s1, d);
              s2 = fma2(d, d, s2);
              float da, db;
              unpack2(d, da, db);
              w[j] = pack_bf16x2(da, db);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(brow + ((static_cast<uint32_t>(j) ^ sw64) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
          if (ci == kChunks - 1) {
            float a, b, q, r;
            unpack2(s1, a, b);
            unpack2(s2, q, r);
            const float sum = a + b;
            const float mean = fmaf(sum, 1.0f / kStatCols, shift);
            const float m2 = fmaxf((q + r) - sum * sum * (1.0f / kStatCols), 0.f);
            if (valid) p.stats_out[row * parts + (n0 + c) / kStatCols] = make_float2(mean, m2);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_out, xbuf0 + s * kFwdXBytes, n0 + c, row0);
            ptx::tma_store_2d(&tmap_aux, bbuf0 + s * kFwdBBytes, n0 + c, row0);
            ptx::tma_store_commit();
            ptx::tma_store_wait_read<1>();        // the previous chunk's stores have left their slot ...
            issue_load(g + kFwdSlots - 1);        // ... which is the one chunk g + kFwdSlots - 1 uses
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      if (lane == 0) ptx::tma_store_wait<0>();
    } else {
      const uint32_t stg = staging_base + static_cast<uint32_t>(warp_idx) * (C::kStagingBufs * 32u * 128u);
      uint32_t stg_buf = 0;
      LnRowStats ln_next;  // forwarded LayerNorm: the row statistics of the NEXT tile, loaded one tile ahead
      auto ln_prefetch = [&](int64_t it) {
        int mb, nb;
        if (pair_tile(it, pair_idx, pair_stride, p.num_m_blocks, p.num_n_blocks, mb, nb))
          ln_stats_load(p, static_cast<int64_t>(mb) * (2 * kBlockM) + static_cast<int>(cta_rank) * kBlockM + quarter * 32 + lane,
                        ln_next);
      };
      if constexpr (ET::kLnApply) ln_prefetch(0);
      for (int64_t it = 0; pair_tile(it, pair_idx, pair_stride, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk); ++it) {
        const int row0 = m_blk * (2 * kBlockM) + static_cast<int>(cta_rank) * kBlockM + quarter * 32;
        const int n0 = n_blk * BLOCK_N;
        float ln_a = 1.f;
        if constexpr (ET::kLnApply) {
          ln_stats_finish(p, ln_next, ln_a);
          ln_prefetch(it + 1);
        }
        ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
        const uint32_t leader_empty = ptx::mapa(tmem_empty_bar(acc), 0);
        epilogue_tile<EPI, C::kStagingBufs, (EPI_WARPS == 16 ? 16 : 32)>(p, &tmap_out, taddr, row0, lane, n0, col_part * kColsPerWarp,
                                            (col_part + 1) * kColsPerWarp, stg, stg_buf, ln_a,
                           [&]() {
                             __syncwarp();
                             if (lane == 0) ptx::mbar_arrive_cluster(leader_empty);
                           });
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      if constexpr (ET::kStaged) {
        if (lane == 0) ptx::tma_store_wait<0>();
      }
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

// Tensor maps are a pure function of (base, shape, leading dim, box, element size, swizzle): the model calls
// duo_gemm with the same few dozen operand descriptions every step (workspace views, packed weights), so the
// encoded maps are kept in a small per-thread cache and cuTensorMapEncodeTiled stays off the launch path.
struct TmapKey {
  const void* base;
  int64_t rows, cols, ld;
  int32_t box_rows, elem_bytes, swizzle_bytes;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           elem_bytes == o.elem_bytes && swizzle_bytes == o.swizzle_bytes;
  }
};
constexpr int kTmapCacheSize = 128;
struct TmapCache {
  TmapKey key[kTmapCacheSize];
  CUtensorMap map[kTmapCacheSize];
  int used = 0, next = 0;
};
thread_local TmapCache g_tmap_cache;

// Row-major [rows, cols] matrix (leading dim ld elements); box = box_rows x (swizzle_bytes of columns),
// swizzle 128 B or 64 B.  elem_bytes 2 -> bf16 (fp16 operands share the encoding), 4 -> fp32.
int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld,
              int box_rows, int elem_bytes, int swizzle_bytes = 128) {
  TmapCache& tc = g_tmap_cache;
  const TmapKey k{base, rows, cols, ld, box_rows, elem_bytes, swizzle_bytes};
  for (int i = 0; i < tc.used; ++i) {
    if (tc.key[i] == k) {
      *tm = tc.map[i];
      return DUO_OK;
    }
  }
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DUO_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(swizzle_bytes / elem_bytes), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
    return DUO_ERR_CUDA;
  }
  const int slot = tc.used < kTmapCacheSize ? tc.used++ : (tc.next = (tc.next + 1) % kTmapCacheSize);
  tc.key[slot] = k;
  tc.map[slot] = *tm;
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
    case kEpiBf16Ln: return launch<BLOCK_N, kEpiBf16Ln>(ta, tb, to, p, st);
    case kEpiGeluBf16Ln: return launch<BLOCK_N, kEpiGeluBf16Ln>(ta, tb, to, p, st);
    case kEpiResidualTma: return launch<BLOCK_N, kEpiResidualTma>(ta, tb, to, p, st);
    case DUO_EPI_SCATTER_F32: return launch<BLOCK_N, DUO_EPI_SCATTER_F32>(ta, tb, to, p, st);
    case DUO_EPI_F32: return launch<BLOCK_N, DUO_EPI_F32>(ta, tb, to, p, st);
    case DUO_EPI_SPLIT_BF16: return launch<BLOCK_N, DUO_EPI_SPLIT_BF16>(ta, tb, to, p, st);
    case DUO_EPI_GELU_SPLIT_BF16: return launch<BLOCK_N, DUO_EPI_GELU_SPLIT_BF16>(ta, tb, to, p, st);
    default: set_error("duo_gemm: unknown epilogue %d", epi); return DUO_ERR_INVALID;
  }
}

template <int EPI, int EPI_WARPS>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                const GemmParams& p, cudaStream_t st) {
  using C = PairCfg<EPI, EPI_WARPS>;
  static uint64_t configured = 0;  // per device
  auto kfn = gemm_tcgen05_pair_kernel<EPI, EPI_WARPS>;
  if (first_use_on_device(configured))
    DUO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(C::kSmemBytes)));
  const int64_t units = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;  // tiles, round-robin over pairs
  const int pairs_max = device_sm_count() / 2;
  const int pairs = static_cast<int>(units < pairs_max ? units : pairs_max);
  kfn<<<2 * pairs, C::kThreads, C::kSmemBytes, st>>>(ta, tb, to, tx, p);
  DUO_LAUNCH_CHECK("gemm_tcgen05_pair_kernel");
  return DUO_OK;
}

// Epilogue warps of the pair kernel: 8 for the GELU epilogues and the token scatter, 4 otherwise.
int dispatch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                  const GemmParams& p, int epi, cudaStream_t st) {
  switch (epi) {
    case DUO_EPI_BF16: return launch_pair<DUO_EPI_BF16, 4>(ta, tb, to, tx, p, st);
#ifndef DUO_GELU_WARPS
#define DUO_GELU_WARPS 8  // 16 (four warps per scheduler, 16-column TMEM loads) measured 2 % slower: csrc/Makefile `tuning`
#endif
    case DUO_EPI_GELU_BF16: return launch_pair<DUO_EPI_GELU_BF16, DUO_GELU_WARPS>(ta, tb, to, tx, p, st);
    case kEpiBf16Ln: return launch_pair<kEpiBf16Ln, 4>(ta, tb, to, tx, p, st);
    case kEpiGeluBf16Ln: return launch_pair<kEpiGeluBf16Ln, DUO_GELU_WARPS>(ta, tb, to, tx, p, st);
    case kEpiResidualTma: return launch_pair<kEpiResidualTma, 4>(ta, tb, to, tx, p, st);
    case kEpiResidualFwd: return launch_pair<kEpiResidualFwd, 4>(ta, tb, to, tx, p, st);
    case kEpiResidualFwdLongK: return launch_pair<kEpiResidualFwdLongK, 4>(ta, tb, to, tx, p, st);
    case DUO_EPI_SCATTER_F32: return launch_pair<DUO_EPI_SCATTER_F32, 8>(ta, tb, to, tx, p, st);
    case DUO_EPI_F32: return launch_pair<DUO_EPI_F32, 4>(ta, tb, to, tx, p, st);
    case DUO_EPI_SPLIT_BF16: return launch_pair<DUO_EPI_SPLIT_BF16, 4>(ta, tb, to, tx, p, st);
    case DUO_EPI_GELU_SPLIT_BF16: return launch_pair<DUO_EPI_GELU_SPLIT_BF16, 8>(ta, tb, to, tx, p, st);
    default: set_error("duo_gemm: unknown epilogue %d", epi); return DUO_ERR_INVALID;
  }
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
  int epi = a->epilogue;
  // statistics forwarding, producer side: residual update that also emits the bf16 copy + row statistics
  const bool fwd = a->xb_out != nullptr || a->stats_out != nullptr;
  DUO_CHECK_ARG(fwd || a->shift_stats == nullptr, "duo_gemm: shift_stats only applies to statistics forwarding (xb_out / stats_out)");
  if (fwd) {
    DUO_CHECK_ARG(epi == DUO_EPI_RESIDUAL_F32 && a->split3 == 0 && a->xb_out && a->stats_out,
                  "duo_gemm: statistics forwarding needs the plain-bf16 residual epilogue and both xb_out and stats_out");
    DUO_CHECK_ARG(a->N % kPairBlockN == 0, "duo_gemm: statistics forwarding needs N %% 256 == 0 (N=%d)", a->N);
    DUO_CHECK_ARG(a->shift_stats != a->stats_out, "duo_gemm: shift_stats must not alias stats_out (other CTAs still read it)");
    DUO_CHECK_ARG((reinterpret_cast<uintptr_t>(a->xb_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->stats_out) & 7) == 0,
                  "duo_gemm: xb_out must be 16-byte aligned, stats_out 8-byte aligned");
  }
  // consumer side: LayerNorm applied in the epilogue from forwarded statistics
  const bool ln_apply = a->ln_stats != nullptr;
  if (ln_apply) {
    DUO_CHECK_ARG((epi == DUO_EPI_BF16 || epi == DUO_EPI_GELU_BF16) && a->split3 == 0 && !a->relu,
                  "duo_gemm: a forwarded LayerNorm needs the BF16 / GELU_BF16 epilogue and plain bf16 operands");
    DUO_CHECK_ARG(a->K % kStatCols == 0 && a->K <= kStatCols * kMaxStatParts, "duo_gemm: a forwarded LayerNorm needs K %% 256 == 0, K <= 1024 (K=%d)", a->K);
    DUO_CHECK_ARG((reinterpret_cast<uintptr_t>(a->ln_stats) & 7) == 0, "duo_gemm: ln_stats must be 8-byte aligned");
    epi = epi == DUO_EPI_BF16 ? kEpiBf16Ln : kEpiGeluBf16Ln;
  }
#ifndef DUO_FWD_LONGK_FROM
#define DUO_FWD_LONGK_FROM 1536  // K at which the long-K forwarding variant takes over (tools/bench_kernels.py --lib sweeps it)
#endif
  if (epi == DUO_EPI_RESIDUAL_F32) epi = !fwd ? kEpiResidualTma : (a->K >= DUO_FWD_LONGK_FROM ? kEpiResidualFwdLongK : kEpiResidualFwd);

  // Tile shape: CTA-pair 256x256 when N allows and there are at least two waves of pair tiles;
  // else 128 x 256 (enough tiles to fill the machine) or 128 x 128.  The forwarding epilogue exists in the
  // pair kernel only.
  int64_t m_blocks = (a->M + kBlockM - 1) / kBlockM;
  int block_n = 128;
  if (a->N % 256 == 0 && m_blocks * (a->N / 256) >= 2 * device_sm_count()) block_n = 256;
  const int64_t m2_blocks = (a->M + 2 * kBlockM - 1) / (2 * kBlockM);
  const bool use_pair = fwd || (a->N % 256 == 0 && m2_blocks * (a->N / 256) >= device_sm_count());
  if (use_pair) {
    block_n = 128;  // W box rows: each CTA loads half of the 256-wide tile
    m_blocks = m2_blocks;
  }

  CUtensorMap ta, tb, to, tx;
  int rc = make_tmap(&ta, a->A, a->M, acols, a->lda, kBlockM, 2);
  if (rc != DUO_OK) return rc;
  rc = make_tmap(&tb, a->W, a->N, kcols, a->ldw, block_n, 2);
  if (rc != DUO_OK) return rc;
  if (epi == DUO_EPI_BF16 || epi == DUO_EPI_GELU_BF16 || epi == kEpiBf16Ln || epi == kEpiGeluBf16Ln) {
    rc = make_tmap(&to, a->out, a->M, a->N, a->ldo, 32, 2);
  } else if (epi == kEpiResidualTma || fwd) {
    rc = make_tmap(&to, a->out, a->M, a->N, a->ldo, 32, 4);
  } else {
    to = ta;  // unused by the direct-store epilogues
  }
  if (rc != DUO_OK) return rc;
  if (fwd) {
    rc = make_tmap(&tx, a->xb_out, a->M, a->N, a->N, 32, 2, 64);
    if (rc != DUO_OK) return rc;
  } else {
    tx = ta;  // unused
  }
  GemmParams p;
  p.stats_out = reinterpret_cast<float2*>(a->stats_out);
  p.shift_stats = reinterpret_cast<const float2*>(a->shift_stats);
  p.ln_stats = reinterpret_cast<const float2*>(a->ln_stats);
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
  if (use_pair) return dispatch_pair(ta, tb, to, tx, p, epi, st);
  return block_n == 256 ? dispatch_epi<256>(ta, tb, to, p, epi, st) : dispatch_epi<128>(ta, tb, to, p, epi, st);
}
