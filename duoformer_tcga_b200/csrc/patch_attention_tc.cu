// Global ("patch" / "region") attention on tcgen05 + TMEM in split-bf16 precision, for groups of at
// most 64 tokens and head_dim 64: the N = P + 1 = 50 tokens of a 224 x 224 tile
// (scale_attention.py:195-207, multiscale_attn.py:205-216: softmax(q k^T * scale) v per (image, head)).
//
// The patch blocks have no residual stream, so their operand rounding compounds (precision
// policy): q, k, v arrive as hi | lo bf16 pairs (x = hi + lo to ~16 mantissa bits, the SPLIT epilogue
// of the QKV GEMM), and both products run as three bf16 UMMAs accumulated in TMEM,
//     S = Qh Kh^T + Qh Kl^T + Ql Kh^T            O = Ph Vh + Ph Vl + Pl Vh,
// which reproduces the fp32 result to ~2^-16.  Same structure as scale_attention_tc.cu: one persistent
// CTA of four warps walks (image, head) problems, TWO CTAs per SM (96 KB each: the kernel is a latency chain per
// problem, round 1 ran one CTA per SM with double-buffered operands); warp 3 issues TMA (six head slices per problem;
// the next problem's loads start once Q K^T has completed and both V halves are transposed) and the MMAs and
// transposes Vh, warp 2 transposes Vl, warps 0-1 own one query row
// per thread: scores from TMEM, softmax in registers, P split into hi | lo in shared memory, O from
// TMEM scaled by 1 / sum and written as a hi | lo pair for the split proj GEMM.
// Rows >= N of the M = 128 operands are whatever follows in shared memory (they feed accumulator
// rows nobody reads); key padding (N..63): P columns are zeros, V padding rows are zeroed once.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace duo {
namespace {

constexpr int kDh = 64;
constexpr int kKeys = 64;                         // padded group size
constexpr uint32_t kTile = kKeys * 128;           // one head slice: 64 rows x 128 B = 8 KB
constexpr uint32_t kBuf = 6 * kTile;              // Qh Ql Kh Kl Vh Vl
constexpr uint32_t kPTile = 128 * 128;            // P hi / lo: M = 128 rows x 64 keys
constexpr uint32_t kSmemData = kBuf + 2 * kPTile + 2 * kTile;  // Q K V (hi, lo) | P hi, lo | V^T hi, lo = 96 KB
constexpr uint32_t kSmemBytes = kSmemData + 64;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0, kColO = 128;

__device__ __forceinline__ void ldmatrix_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// V[key][d] (TMA layout, 128B swizzle) -> V^T[d][key] (K-major, 128B swizzle), 64 x 64, one warp.
__device__ __forceinline__ void transpose_v(uint32_t v_tile, uint32_t vt_tile, int lane) {
#pragma unroll 2
  for (int kb = 0; kb < kKeys / 8; ++kb) {
    const int key = 8 * kb + (lane & 7);
#pragma unroll
    for (int cq = 0; cq < 2; ++cq) {
      const int chunk = 4 * cq + (lane >> 3);
      uint32_t r[4];
      ldmatrix_x4_t(v_tile + static_cast<uint32_t>(key * 128 + ((chunk ^ (key & 7)) << 4)), r[0], r[1], r[2], r[3]);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int d = 8 * (4 * cq + m) + (lane >> 2);
        const uint32_t dst = vt_tile + static_cast<uint32_t>(d * 128 + ((kb ^ (d & 7)) << 4) + 4 * (lane & 3));
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(r[m]) : "memory");
      }
    }
  }
}

__global__ void __launch_bounds__(128, 2)
patch_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ out,
                          int N, int H, int64_t problems, float scale_log2e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  const uint32_t p_base = base + kBuf;              // P hi, P lo
  const uint32_t vt_base = p_base + 2 * kPTile;     // V^T hi, V^T lo
  const uint32_t bar_base = base + kSmemData;
  const uint32_t full_bar0 = bar_base, s_full = bar_base + 16, p_ready = bar_base + 24, o_full = bar_base + 32;
  const uint32_t vt_ready = bar_base + 40;
  const uint32_t tmem_slot = bar_base + 48;
  uint32_t* tmem_slot_generic = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = H * kDh;
  const uint32_t load_bytes = 6u * static_cast<uint32_t>(N) * 128u;
  const int nk = (N + 15) >> 4;

  // V padding rows (hi and lo): zero once
  for (int i = threadIdx.x; i < 2 * (kKeys - N) * 8; i += 128) {
    const int t = i / ((kKeys - N) * 8);  // hi / lo
    const int j = i - t * (kKeys - N) * 8;
    const uint32_t dst = base + (4 + t) * kTile + static_cast<uint32_t>((N + (j >> 3)) * 128 + ((j & 7) << 4));
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
  }
  if (warp == 3) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_qkv);
      ptx::mbar_init(full_bar0, 1);
      ptx::mbar_init(s_full, 1);
      ptx::mbar_init(p_ready, 2);   // one arrival per softmax warp
      ptx::mbar_init(o_full, 1);
      ptx::mbar_init(vt_ready, 1);  // warp 2: V^T lo written
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

  // qkv row = [hi: q k v | lo: q k v], each 3 * D wide
  auto issue_loads = [&](int64_t prob) {  // warp 3, lane 0
    const int64_t g = prob / H;
    const int h = static_cast<int>(prob - g * H);
    const uint32_t dst = base;
    const uint32_t bar = full_bar0;
    const int32_t row = static_cast<int32_t>(g * N);
    ptx::mbar_arrive_expect_tx(bar, load_bytes);
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      ptx::tma_load_2d(dst + (2 * which) * kTile, &tmap_qkv, bar, which * D + h * kDh, row);              // hi
      ptx::tma_load_2d(dst + (2 * which + 1) * kTile, &tmap_qkv, bar, 3 * D + which * D + h * kDh, row);  // lo
    }
  };

  const int64_t first = blockIdx.x;
  const int64_t stride = gridDim.x;
  constexpr uint32_t idesc = ptx::make_idesc_bf16(128, kKeys);

  if (warp == 3) {
    // ===================== control warp: TMA, MMA issue, V^T hi =====================
    if (lane == 0 && first < problems) issue_loads(first);
    int it = 0;
    for (int64_t prob = first; prob < problems; prob += stride, ++it) {
      const uint32_t buf = base;
      ptx::mbar_wait(full_bar0, static_cast<uint32_t>(it & 1));
      if (lane == 0) {
        ptx::tc_fence_after();
        const uint64_t dqh = ptx::make_smem_desc_sw128(buf), dql = ptx::make_smem_desc_sw128(buf + kTile);
        const uint64_t dkh = ptx::make_smem_desc_sw128(buf + 2 * kTile), dkl = ptx::make_smem_desc_sw128(buf + 3 * kTile);
        uint32_t accumulate = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {  // Qh Kh, Qh Kl, Ql Kh
          const uint64_t da = pass == 2 ? dql : dqh;
          const uint64_t db = pass == 1 ? dkl : dkh;
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k) {
            ptx::umma_bf16(tmem_base + kColS, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                           accumulate);
            accumulate = 1;
          }
        }
        ptx::umma_commit(s_full);
      }
      if (it > 0) ptx::mbar_wait(o_full, static_cast<uint32_t>((it - 1) & 1));  // previous P V has consumed V^T
      transpose_v(buf + 4 * kTile, vt_base, lane);
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_wait(vt_ready, static_cast<uint32_t>(it & 1));  // V^T lo (warp 2)
        if (prob + stride < problems) {
          ptx::mbar_wait(s_full, static_cast<uint32_t>(it & 1));  // Q / K consumed, both V halves transposed
          issue_loads(prob + stride);
        }
        ptx::mbar_wait(p_ready, static_cast<uint32_t>(it & 1));
        ptx::tc_fence_after();
        const uint64_t dph = ptx::make_smem_desc_sw128(p_base), dpl = ptx::make_smem_desc_sw128(p_base + kPTile);
        const uint64_t dvh = ptx::make_smem_desc_sw128(vt_base), dvl = ptx::make_smem_desc_sw128(vt_base + kTile);
        uint32_t accumulate = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {  // Ph Vh, Ph Vl, Pl Vh
          const uint64_t da = pass == 2 ? dpl : dph;
          const uint64_t db = pass == 1 ? dvl : dvh;
          for (int k = 0; k < nk; ++k) {
            ptx::umma_bf16(tmem_base + kColO, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                           accumulate);
            accumulate = 1;
          }
        }
        ptx::umma_commit(o_full);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===================== helper warp: V^T lo =====================
    int it = 0;
    for (int64_t prob = first; prob < problems; prob += stride, ++it) {
      ptx::mbar_wait(full_bar0, static_cast<uint32_t>(it & 1));
      if (it > 0) ptx::mbar_wait(o_full, static_cast<uint32_t>((it - 1) & 1));
      transpose_v(base + 5 * kTile, vt_base + kTile, lane);
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(vt_ready);
    }
  } else {
    // ===================== softmax / output warps 0-1: thread = query row (0..63) =====================
    const int r = warp * 32 + lane;
    const uint32_t x7 = static_cast<uint32_t>(r & 7);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    int it = 0;
    for (int64_t prob = first; prob < problems; prob += stride, ++it) {
      const int64_t g = prob / H;
      const int h = static_cast<int>(prob - g * H);
      ptx::mbar_wait(s_full, static_cast<uint32_t>(it & 1));
      ptx::tc_fence_after();
      uint32_t v0[32], v1[32];
      ptx::tmem_ld_32x32(taddr + kColS, v0);
      ptx::tmem_ld_32x32(taddr + kColS + 32, v1);
      ptx::tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < N) mx = fmaxf(mx, __uint_as_float(v0[j]));
        if (32 + j < N) mx = fmaxf(mx, __uint_as_float(v1[j]));
      }
      const float off = mx * scale_log2e;
      float sum = 0.f;
      auto emit = [&](const uint32_t (&v)[32], int key0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int key = key0 + 8 * c + 2 * j;
            const float pa = key < N ? exp2f(fmaf(__uint_as_float(v[8 * c + 2 * j]), scale_log2e, -off)) : 0.f;
            const float pb = key + 1 < N ? exp2f(fmaf(__uint_as_float(v[8 * c + 2 * j + 1]), scale_log2e, -off)) : 0.f;
            sum += pa + pb;
            pack_split2(pa, pb, hi[j], lo[j]);
          }
          const uint32_t off_b = static_cast<uint32_t>(r) * 128u + ((static_cast<uint32_t>((key0 >> 3) + c) ^ x7) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_base + off_b), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                       "r"(hi[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_base + kPTile + off_b), "r"(lo[0]), "r"(lo[1]),
                       "r"(lo[2]), "r"(lo[3])
                       : "memory");
        }
      };
      emit(v0, 0);
      emit(v1, 32);
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
      if (r < N) {
        const float inv = 1.0f / sum;
        __nv_bfloat16* orow = out + (g * N + r) * (2 * static_cast<int64_t>(D)) + h * kDh;  // hi | lo halves, D apart
        uint4* oh = reinterpret_cast<uint4*>(orow);
        uint4* ol = reinterpret_cast<uint4*>(orow + D);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t(&v)[32] = c < 4 ? v0 : v1;
          const int e = 8 * (c & 3);
          uint4 hv, lv;
          pack_split2(__uint_as_float(v[e + 0]) * inv, __uint_as_float(v[e + 1]) * inv, hv.x, lv.x);
          pack_split2(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv, hv.y, lv.y);
          pack_split2(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv, hv.z, lv.z);
          pack_split2(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv, hv.w, lv.w);
          oh[c] = hv;
          ol[c] = lv;
        }
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

// qkv split bf16 [groups * N, 2 * 3 * H * 64] (hi | lo) -> out split bf16 [groups * N, 2 * H * 64]; N <= 64.
int launch_patch_attention_tc(const void* qkv, void* out, int64_t groups, int N, int H, float scale, cudaStream_t st) {
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
  const int64_t rows = groups * N;
  const int64_t cols = 6LL * H * kDh;
  if (rows >= (int64_t(1) << 31)) {
    set_error("duo_group_attention: too many rows for the tcgen05 kernel");
    return DUO_ERR_INVALID;
  }
  CUtensorMap tm;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kDh), static_cast<cuuint32_t>(N)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for the split qkv tensor", static_cast<int>(r));
    return DUO_ERR_CUDA;
  }
  static uint64_t configured = 0;  // per device
  if (first_use_on_device(configured))
    DUO_CUDA(cudaFuncSetAttribute(patch_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kSmemBytes)));
  const int64_t problems = groups * H;
  const int64_t max_ctas = 2LL * device_sm_count();
  const unsigned grid = static_cast<unsigned>(problems < max_ctas ? problems : max_ctas);
  patch_attention_tc_kernel<<<grid, 128, kSmemBytes, st>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), N, H, problems,
                                                           scale * 1.4426950408889634f);
  DUO_LAUNCH_CHECK("patch_attention_tc_kernel");
  return DUO_OK;
}

}  // namespace duo
