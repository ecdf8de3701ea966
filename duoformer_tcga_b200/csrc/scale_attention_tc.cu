// Scale attention on the 5th-generation tensor cores (tcgen05 + TMEM) for 64 < S <= 96 tokens per
// group and head_dim 64 — the S = 86 case of the 4-scale model (scale_attention.py:28-45,
// multiscale_attn.py:149-166: softmax(q k^T * scale) v inside every (patch, head)).
//
// One persistent CTA of four warps walks (group, head) problems; THREE CTAs share an SM (74 KB of shared memory and
// 128 TMEM columns each: a problem is a ~3 000-cycle dependent chain — MMA, TMEM round trips, a MUFU-bound softmax,
// barriers — so the kernel lives on problems in flight; with two CTAs per SM it ran at 0.94 of the HBM peak alone at
// burst clocks but 0.60 inside the power-capped step at ~1.05 GHz).
//   warp 3   : lane 0 issues TMA loads (Q, K, V head slices, one box of S rows x 128 B each; the next problem's loads
//              start as soon as Q K^T has completed and V is transposed) and all tcgen05.mma; the whole warp
//              transposes V into the K-major V^T operand (ldmatrix.trans -> st.shared) while the scores are computed.
//   warps 0-2: thread = query row.  scores from TMEM (tcgen05.ld) -> softmax in registers (no
//              shuffles) -> P (bf16) into shared memory -> after the second MMA, O from TMEM,
//              scaled by 1/sum, transposed through the warp's (dead) P rows and stored as full
//              128-byte lines.
// (Round 1 / early round 2: two CTAs per SM with double-buffered Q | K | V, S and O in separate TMEM columns.)
// (Measured and dropped in round 2: two threads per query row — eight warps per CTA splitting the score / output
// columns, a named barrier per row quarter for the maximum, a helper warp for half of the V transpose: 0.303 ms per 64
// images against 0.270 alone, 22.4 against 17.8 ms per step.)
//   MMA 1    : S[128 x 96] = Q[128 x 64] K^T          4 x UMMA 128x96x16, accumulator columns [0, 96)
//   MMA 2    : O[128 x 64] = P[128 x 96] V^T^T        ceil(S/16) x UMMA 128x64x16, columns [0, 64): O ALIASES S (every
//              score has been read when P is complete; the next problem's MMA 1 waits until O has been read: `o_read`)
// Q and K tiles sit 88 rows apart (S <= 88: the model's 86; otherwise 96): rows >= S of the M = 128 / N = 96 operands are
// whatever follows them in shared memory (K rows for Q, V rows for K): they only feed accumulator rows that are never read
// and score columns that are masked.  Key padding (S..95): P columns are written as zeros and the V padding rows are
// zeroed once (TMA boxes never touch them), so no NaN can enter a real row.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace duo {
namespace {

constexpr int kDh = 64;
constexpr int kKeysPad = 96;
constexpr uint32_t kTile = kKeysPad * 128;       // a 96-row tile: V, and each half of P
constexpr uint32_t kVtBlock = kDh * 128;         // V^T: 64 rows x (64 keys) per block
constexpr uint32_t kPBytes = 2 * kTile;          // P: keys 0..63 | keys 64..95, 96 rows x 128 B each
// Q | K | V with Q and K `qk_rows` rows apart (a multiple of 8: 1024-byte aligned swizzle atoms)
__host__ __device__ constexpr uint32_t qk_pitch(int S) { return static_cast<uint32_t>(S <= 88 ? 88 : 96) * 128u; }
__host__ __device__ constexpr uint32_t smem_data(int S) { return 2 * qk_pitch(S) + kTile + kPBytes + 2 * kVtBlock; }
__host__ __device__ constexpr uint32_t smem_bytes(int S) { return smem_data(S) + 64; }  // S <= 88: 74 KB + barriers, three CTAs per SM
constexpr uint32_t kTmemCols = 128;
constexpr uint32_t kColS = 0, kColO = 0;

__device__ __forceinline__ void ldmatrix_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// S_CT > 0: S known at compile time (the model's 86); 0: runtime S.
template <int S_CT>
__global__ void __launch_bounds__(128, 3)
scale_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ out,
                          int S_rt, int H, int64_t problems, float scale_log2e) {
  const int S = S_CT > 0 ? S_CT : S_rt;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();  // the 128-byte swizzle patterns of TMA and UMMA assume 1024-byte aligned tiles
  const uint32_t kQK = qk_pitch(S);
  const uint32_t k_tile = base + kQK, v_tile = base + 2 * kQK;
  const uint32_t p_base = v_tile + kTile;
  const uint32_t vt_base = p_base + kPBytes;
  const uint32_t bar_base = base + smem_data(S);
  const uint32_t full_bar0 = bar_base, o_read = bar_base + 8, s_full = bar_base + 16, p_ready = bar_base + 24, o_full = bar_base + 32;
  const uint32_t tmem_slot = bar_base + 40;
  uint32_t* tmem_slot_generic = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = H * kDh;
  const uint32_t load_bytes = 3u * static_cast<uint32_t>(S) * 128u;
  const int nk = (S + 15) >> 4;  // 16-key steps of the second MMA

  // V padding rows: zero once
  for (int j = threadIdx.x; j < (kKeysPad - S) * 8; j += 128) {
    const uint32_t dst = v_tile + static_cast<uint32_t>((S + (j >> 3)) * 128 + ((j & 7) << 4));
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
  }
  if (warp == 3) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_qkv);
      ptx::mbar_init(full_bar0, 1);
      ptx::mbar_init(o_read, 3);   // one arrival per softmax warp: O (= the S columns) has been read
      ptx::mbar_init(s_full, 1);
      ptx::mbar_init(p_ready, 3);  // one arrival per softmax warp
      ptx::mbar_init(o_full, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<kTmemCols>(tmem_slot);
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_generic;

  auto issue_loads = [&](int64_t prob) {  // warp 3, lane 0
    const int64_t g = prob / H;
    const int h = static_cast<int>(prob - g * H);
    const int32_t row = static_cast<int32_t>(g * S);
    ptx::mbar_arrive_expect_tx(full_bar0, load_bytes);
    ptx::tma_load_2d(base, &tmap_qkv, full_bar0, h * kDh, row);
    ptx::tma_load_2d(k_tile, &tmap_qkv, full_bar0, D + h * kDh, row);
    ptx::tma_load_2d(v_tile, &tmap_qkv, full_bar0, 2 * D + h * kDh, row);
  };

  const int64_t first = blockIdx.x;
  const int64_t stride = gridDim.x;

  if (warp == 3) {
    // ===================== control warp: TMA, MMA issue, V transpose =====================
    constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, kKeysPad);
    constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, kDh);
    if (lane == 0 && first < problems) issue_loads(first);
    int it = 0;
    for (int64_t prob = first; prob < problems; prob += stride, ++it) {
      ptx::mbar_wait(full_bar0, static_cast<uint32_t>(it & 1));
      // the previous problem's O sits in the score columns: wait until the softmax warps have read it
      if (it > 0) ptx::mbar_wait(o_read, static_cast<uint32_t>((it - 1) & 1));
      if (lane == 0) {
        ptx::tc_fence_after();
        const uint64_t dq = ptx::make_smem_desc_sw128(base);
        const uint64_t dk = ptx::make_smem_desc_sw128(k_tile);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          ptx::umma_bf16(tmem_base + kColS, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k),
                         idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
      }
      // the previous problem's second MMA has consumed V^T
      if (it > 0) ptx::mbar_wait(o_full, static_cast<uint32_t>((it - 1) & 1));
      // ---- V[key][d] -> V^T[d][key] (K-major, 128B swizzle, two blocks of 64 keys) ----
      {
#pragma unroll 2
        for (int kb = 0; kb < kKeysPad / 8; ++kb) {
          const int key = 8 * kb + (lane & 7);
          const uint32_t dst_blk = vt_base + static_cast<uint32_t>(kb >> 3) * kVtBlock;
          const int kc = kb & 7;
#pragma unroll
          for (int cq = 0; cq < 2; ++cq) {
            const int chunk = 4 * cq + (lane >> 3);
            uint32_t r[4];
            ldmatrix_x4_t(v_tile + static_cast<uint32_t>(key * 128 + ((chunk ^ (key & 7)) << 4)), r[0], r[1], r[2], r[3]);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int d = 8 * (4 * cq + m) + (lane >> 2);
              const uint32_t dst = dst_blk + static_cast<uint32_t>(d * 128 + ((kc ^ (d & 7)) << 4) + 4 * (lane & 3));
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(r[m]) : "memory");
            }
          }
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        // Q | K | V are free once Q K^T has completed (V is already transposed): fetch the next problem
        if (prob + stride < problems) {
          ptx::mbar_wait(s_full, static_cast<uint32_t>(it & 1));
          issue_loads(prob + stride);
        }
        ptx::mbar_wait(p_ready, static_cast<uint32_t>(it & 1));
        ptx::tc_fence_after();
        const uint64_t dp0 = ptx::make_smem_desc_sw128(p_base);           // P keys 0..63
        const uint64_t dp1 = ptx::make_smem_desc_sw128(p_base + kTile);   // P keys 64..95
        const uint64_t dv0 = ptx::make_smem_desc_sw128(vt_base);
        const uint64_t dv1 = ptx::make_smem_desc_sw128(vt_base + kVtBlock);
        for (int k = 0; k < nk; ++k) {
          const uint64_t da = (k < 4 ? dp0 : dp1) + static_cast<uint64_t>(2 * (k & 3));
          const uint64_t db = (k < 4 ? dv0 : dv1) + static_cast<uint64_t>(2 * (k & 3));
          ptx::umma_bf16(tmem_base + kColO, da, db, idesc_o, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / output warps: thread = query row =====================
    const int r = warp * 32 + lane;
    const uint32_t x7 = static_cast<uint32_t>(r & 7);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    int it = 0;
    for (int64_t prob = first; prob < problems; prob += stride, ++it) {
      const int64_t g = prob / H;
      const int h = static_cast<int>(prob - g * H);
      ptx::mbar_wait(s_full, static_cast<uint32_t>(it & 1));
      ptx::tc_fence_after();
      uint32_t v0[32], v1[32], v2[32];
      ptx::tmem_ld_32x32(taddr + kColS, v0);
      ptx::tmem_ld_32x32(taddr + kColS + 32, v1);
      ptx::tmem_ld_32x32(taddr + kColS + 64, v2);
      ptx::tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (64 + j < S) mx = fmaxf(mx, __uint_as_float(v2[j]));
      const float off = mx * scale_log2e;
      // probabilities, 8 keys (one 16-byte chunk of the P row) at a time; the scaling and the row sum run on packed
      // fp32 pairs (FFMA2 / FADD2: half the issue slots of the scalar form — the kernel is issue / latency bound at
      // the ~1.1 GHz the power-capped step runs at)
      const uint64_t sc2 = pack2(scale_log2e, scale_log2e), noff2 = pack2(-off, -off);
      uint64_t sum2 = pack2(0.f, 0.f);
      auto emit = [&](const uint32_t (&v)[32], int key0, uint32_t tile) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int key = key0 + 8 * c + 2 * j;
            float e0, e1;
            unpack2(fma2(pack2(__uint_as_float(v[8 * c + 2 * j]), __uint_as_float(v[8 * c + 2 * j + 1])), sc2, noff2), e0, e1);
            const float p0 = key < S ? ex2_approx(e0) : 0.f;
            const float p1 = key + 1 < S ? ex2_approx(e1) : 0.f;
            sum2 = add2(sum2, pack2(p0, p1));
            w[j] = pack_bf16x2(p0, p1);
          }
          const uint32_t kc = static_cast<uint32_t>(((key0 & 63) >> 3) + c);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + static_cast<uint32_t>(r) * 128u + ((kc ^ x7) << 4)),
                       "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      };
      emit(v0, 0, p_base);
      emit(v1, 32, p_base);
      emit(v2, 64, p_base + kTile);
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_ready);

      ptx::mbar_wait(o_full, static_cast<uint32_t>(it & 1));
      ptx::tc_fence_after();
      ptx::tmem_ld_32x32(taddr + kColO, v0);
      ptx::tmem_ld_32x32(taddr + kColO + 32, v1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_read);  // the score columns may be overwritten by the next problem's Q K^T
      {
        // O row -> this thread's (dead) P row, then the warp stores four complete 128-byte rows per instruction
        float sum_a, sum_b;
        unpack2(sum2, sum_a, sum_b);
        const float inv = 1.0f / (sum_a + sum_b);
        const uint64_t inv2 = pack2(inv, inv);
        const uint32_t my_row = p_base + static_cast<uint32_t>(r) * 128u;
        auto scaled = [&](const uint32_t (&v)[32], int i) {  // bf16x2 of (v[i], v[i+1]) * inv
          float a, b;
          unpack2(mul2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), inv2), a, b);
          return pack_bf16x2(a, b);
        };
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + ((static_cast<uint32_t>(c) ^ x7) << 4)),
                       "r"(scaled(v0, 8 * c)), "r"(scaled(v0, 8 * c + 2)), "r"(scaled(v0, 8 * c + 4)), "r"(scaled(v0, 8 * c + 6))
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + ((static_cast<uint32_t>(4 + c) ^ x7) << 4)),
                       "r"(scaled(v1, 8 * c)), "r"(scaled(v1, 8 * c + 2)), "r"(scaled(v1, 8 * c + 4)), "r"(scaled(v1, 8 * c + 6))
                       : "memory");
        }
        __syncwarp();
        __nv_bfloat16* obase = out + (g * S) * D + h * kDh + (lane & 7) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = warp * 32 + 4 * i + (lane >> 3);
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                       : "r"(p_base + static_cast<uint32_t>(rr * 128 + (((lane & 7) ^ (rr & 7)) << 4))));
          if (rr < S) *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(rr) * D) = make_uint4(w0, w1, w2, w3);
        }
        __syncwarp();  // rows are free for the next problem's P
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// qkv [groups * S, 3 * H * 64] bf16 -> out [groups * S, H * 64] bf16; 64 < S <= 96.
int launch_scale_attention_tc(const void* qkv, void* out, int64_t groups, int S, int H, float scale, cudaStream_t st) {
  static PFN_encodeTiled encode = nullptr;
  if (encode == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return DUO_ERR_CUDA;
    }
    encode = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  const int64_t rows = groups * S;
  const int64_t cols = 3LL * H * kDh;
  if (rows >= (int64_t(1) << 31)) {
    set_error("duo_group_attention: too many rows for the tcgen05 kernel");
    return DUO_ERR_INVALID;
  }
  CUtensorMap tm;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kDh), static_cast<cuuint32_t>(S)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for the qkv tensor", static_cast<int>(r));
    return DUO_ERR_CUDA;
  }
  static uint64_t configured = 0;  // per device
  if (first_use_on_device(configured)) {
    DUO_CUDA(cudaFuncSetAttribute(scale_attention_tc_kernel<86>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes(86))));
    DUO_CUDA(cudaFuncSetAttribute(scale_attention_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes(96))));
    // the three-CTAs-per-SM design needs the whole 228 KB carve-out
    DUO_CUDA(cudaFuncSetAttribute(scale_attention_tc_kernel<86>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DUO_CUDA(cudaFuncSetAttribute(scale_attention_tc_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  const int64_t problems = groups * H;
  const int64_t max_ctas = 3LL * device_sm_count();
  const unsigned grid = static_cast<unsigned>(problems < max_ctas ? problems : max_ctas);
  if (S == 86)
    scale_attention_tc_kernel<86><<<grid, 128, smem_bytes(86), st>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), S, H, problems,
                                                                     scale * 1.4426950408889634f);
  else
    scale_attention_tc_kernel<0><<<grid, 128, smem_bytes(S), st>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), S, H, problems,
                                                                   scale * 1.4426950408889634f);
  DUO_LAUNCH_CHECK("scale_attention_tc_kernel");
  return DUO_OK;
}

}  // namespace duo
