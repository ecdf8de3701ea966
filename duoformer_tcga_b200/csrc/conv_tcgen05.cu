// Implicit-GEMM convolution on tcgen05 / TMEM for the ResNet-50 trunk (and any NHWC 1x1 / 3x3 convolution with
// 64-multiple channel counts): the trunk taps of model_wo_extra_params.py:214-224 / resnet50ssl.py:35-45 without cuDNN.
//
//   out[b, ho, wo, n] = act( sum_{ky, kx, c} in[b, ho*s + ky - pad, wo*s + kx - pad, c] * W[n, (ky, kx, c)]
//                            + bias[n] (+ residual[b, ho, wo, n]) )
//
// NHWC 16-bit activations (fp16 or bf16), fp32 accumulation in TMEM, BatchNorm folded into W / bias by the host.
//
// No im2col matrix exists anywhere: an M tile is a BOX of 128 output pixels (bw x bh x bb along w, h, batch; all
// powers of two dividing the map, e.g. 8 x 8 x 2 on 56 x 56, 1 x 1 x 128 on 7 x 7) and the K loop walks the filter taps.
// For tap (ky, kx) and channel block c the A operand tile is ONE 4-D TMA box load of the input tensor map
//   dims (C, W, H, B), box (64, bw*s, bh*s, bb), element strides (1, s, s, 1)  at  (c, w0*s + kx - pad, h0*s + ky - pad, b0):
// the TMA unit applies the convolution stride (element strides) and the zero padding (out-of-bounds fill) and writes the
// 128 pixel rows densely, 128 B each, in the SWIZZLE_128B layout the UMMA descriptor expects — the same shared-memory
// tile a dense [128, 64] matrix load produces.  The epilogue (bias, residual add, ReLU, 16-bit pack) stages 32 pixel
// rows x 64 channels per warp and writes them with 4-D TMA box stores (the box clips batch rows past B).
//
// The 7 x 7 / stride 2 stem (3 input channels) runs on the same kernel: duo_stem_pack writes the image as a zero-padded
// NHWC8 tensor [B, H, W + 8, 8] (3 pad pixels left, 5 right, channels 3..7 zero) and the input tensor map describes
// OVERLAPPING windows — dim0 = 64 elements (8 pixels x 8 channels), dim1 = output column with a 2-pixel (32 B) stride —
// so one filter row is one 64-wide K block (7 real taps x 3 real channels, the rest meets zero weights): K = 7 * 64.
//
// Structure as in gemm_tcgen05.cu: persistent CTAs, N fastest; warps 0..3 epilogue (one TMEM lane quarter each),
// warp 4 TMA producer, warp 5 MMA issuer; kStages-deep operand ring, two TMEM accumulator buffers.
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace duo {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr uint32_t kStagingBytesPerWarp = 2 * 32 * 128;  // two 32-row x 128 B buffers

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int kStages = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static constexpr uint32_t kStagingBytes = 4 * kStagingBytesPerWarp;
  static constexpr uint32_t kBarrierBytes = (2 * kStages + 4) * 8 + 8;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
};

struct ConvParams {
  const float* bias;
  const void* residual;  // NHWC [B, Ho, Wo, Cout], same 16-bit type as out, or NULL
  int32_t B, Ho, Wo, Cout;
  int32_t cin_blocks;        // K blocks per filter tap
  int32_t taps_x, taps_y;    // filter window walked by the K loop
  int32_t mul_w, mul_h;      // input coordinate of tap (kx, ky) for output (w, h): w * mul_w + kx + off_w
  int32_t off_w, off_h;
  int32_t bw_log2, bh_log2;  // M tile = 2^bw_log2 x 2^bh_log2 x (128 >> (bw_log2 + bh_log2)) output pixels (w, h, batch)
  int32_t tiles_w, tiles_h;
  int32_t num_m_blocks, num_n_blocks;
  int32_t relu;
  int32_t f16;      // element type of out / residual: 1 fp16, 0 bf16
  int32_t in_f16;   // element type of in / weight (host side: tensor maps, instruction descriptor)
  uint32_t idesc_mask;
};

__device__ __forceinline__ void tile_origin(const ConvParams& p, int m_blk, int& w0, int& h0, int& b0) {
  const int tw = m_blk % p.tiles_w;
  const int t = m_blk / p.tiles_w;
  const int th = t % p.tiles_h;
  const int tb = t / p.tiles_h;
  w0 = tw << p.bw_log2;
  h0 = th << p.bh_log2;
  b0 = tb * (kBlockM >> (p.bw_log2 + p.bh_log2));
}

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t smem_src, int32_t c0, int32_t c1, int32_t c2,
                                             int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ void unpack16x2(uint32_t u, bool f16, float& a, float& b) {
  if (f16) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&u));
    a = t.x;
    b = t.y;
  } else {
    a = __uint_as_float(u << 16);
    b = __uint_as_float(u & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack16x2(float a, float b, bool f16) {
  if (f16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  return pack_bf16x2(a, b);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_out, const ConvParams p) {
  using C = ConvCfg<BLOCK_N>;
  constexpr int kStages = C::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + kStages * C::kStageBytes;
  const uint32_t bar_base = staging_base + C::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  uint32_t* tmem_ptr_generic = reinterpret_cast<uint32_t*>(smem_raw + (tmem_ptr_smem - ptx::smem_u32(smem_raw)));

  constexpr int kTmaWarp = 4, kMmaWarp = 5;
  const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == kTmaWarp && lane == 0) {
    ptx::prefetch_tmap(&tmap_in);
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_out);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar(a), 1);
      ptx::mbar_init(tmem_empty_bar(a), 4);  // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == kMmaWarp) ptx::tmem_alloc<C::kTmemCols>(tmem_ptr_smem);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  const int num_k_blocks = p.taps_x * p.taps_y * p.cin_blocks;
  const int64_t num_tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;

  if (warp_idx == kTmaWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = static_cast<int>(tile / p.num_n_blocks);
        const int n_blk = static_cast<int>(tile - static_cast<int64_t>(m_blk) * p.num_n_blocks);
        int w0, h0, b0;
        tile_origin(p, m_blk, w0, h0, b0);
        const int cw = w0 * p.mul_w + p.off_w;
        const int ch = h0 * p.mul_h + p.off_h;
        int kb = 0;
        for (int ky = 0; ky < p.taps_y; ++ky) {
          for (int kx = 0; kx < p.taps_x; ++kx) {
            for (int cb = 0; cb < p.cin_blocks; ++cb, ++kb) {
              ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t sa = smem_base + stage * C::kStageBytes;
              const uint32_t sb = sa + C::kABytes;
              ptx::mbar_arrive_expect_tx(full_bar(stage), C::kStageBytes);
              tma_load_4d(sa, &tmap_in, full_bar(stage), cb * kBlockK, cw + kx, ch + ky, b0);
              ptx::tma_load_2d(sb, &tmap_w, full_bar(stage), kb * kBlockK, n_blk * BLOCK_N);
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
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
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            ptx::umma_bf16(tmem_d, desc_a + static_cast<uint64_t>(2 * k), desc_b + static_cast<uint64_t>(2 * k), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit(empty_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit(tmem_full_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue warps (0..3) =====================
    const int quarter = warp_idx & 3;
    const uint32_t stg = staging_base + static_cast<uint32_t>(quarter) * kStagingBytesPerWarp;
    const uint32_t my_row_off = static_cast<uint32_t>(lane) * 128u;
    uint32_t stg_buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool f16 = p.f16 != 0;
    const int bwm = (1 << p.bw_log2) - 1, bhm = (1 << p.bh_log2) - 1;
    const int wh_log2 = p.bw_log2 + p.bh_log2;
    const int r = quarter * 32 + lane;  // this thread's pixel inside the tile box
    const int r0 = quarter * 32;        // first pixel of this warp's slab (a sub-box of the tile box)
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = static_cast<int>(tile / p.num_n_blocks);
      const int n_blk = static_cast<int>(tile - static_cast<int64_t>(m_blk) * p.num_n_blocks);
      const int n0 = n_blk * BLOCK_N;
      int w0, h0, b0;
      tile_origin(p, m_blk, w0, h0, b0);
      const int pw = w0 + (r & bwm), ph = h0 + ((r >> p.bw_log2) & bhm), pb = b0 + (r >> wh_log2);
      const int sw0 = w0 + (r0 & bwm), sh0 = h0 + ((r0 >> p.bw_log2) & bhm), sb0 = b0 + (r0 >> wh_log2);
      const bool valid = pb < p.B;
      const uint16_t* res_row = nullptr;
      if (p.residual != nullptr && valid)
        res_row = reinterpret_cast<const uint16_t*>(p.residual) +
                  ((static_cast<int64_t>(pb) * p.Ho + ph) * p.Wo + pw) * p.Cout + n0;
      ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 64) {
        const uint32_t buf = stg + stg_buf * (32u * 128u) + my_row_off;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          float4 bia[8];
          uint4 res[4];
          ptx::tmem_ld_32x32(taddr + static_cast<uint32_t>(c + 32 * h), v);
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c + 32 * h);
#pragma unroll
            for (int j = 0; j < 8; ++j) bia[j] = __ldg(b4 + j);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) bia[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (res_row != nullptr) {
            const uint4* r4 = reinterpret_cast<const uint4*>(res_row + c + 32 * h);
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = __ldg(r4 + j);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = make_uint4(0u, 0u, 0u, 0u);  // +0.0 in either format
          }
          ptx::tmem_ld_wait();
          if (h == 1 && c + 64 >= BLOCK_N) {  // accumulator fully read: hand the TMEM buffer back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(acc));
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bia[j].x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bia[j].y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bia[j].z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bia[j].w;
          }
          if (p.residual != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t u[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float a, b;
                unpack16x2(u[i], f16, a, b);
                f[8 * j + 2 * i] += a;
                f[8 * j + 2 * i + 1] += b;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (h == 0) {
            if (lane == 0) ptx::tma_store_wait_read<1>();  // buffer `stg_buf` no longer being read
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)  // 16-byte chunk (4h + j) of this pixel row, XOR-swizzled
            st_shared_v4(buf + (static_cast<uint32_t>((4 * h + j) ^ (lane & 7)) << 4),
                         pack16x2(f[8 * j + 0], f[8 * j + 1], f16), pack16x2(f[8 * j + 2], f[8 * j + 3], f16),
                         pack16x2(f[8 * j + 4], f[8 * j + 5], f16), pack16x2(f[8 * j + 6], f[8 * j + 7], f16));
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmap_out, stg + stg_buf * (32u * 128u), n0 + c, sw0, sh0, sb0);
          ptx::tma_store_commit();
        }
        stg_buf ^= 1u;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (lane == 0) ptx::tma_store_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// Image -> zero-padded NHWC8 operand of the stem convolution: out[b, y, 3 + x, c] = scale * x[b, c, y, x] for c < 3,
// zero elsewhere ([B, H, W + 8, 8]); any input strides (NCHW or channels-last), one 16-byte store per pixel.
__global__ void stem_pack_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sh, int64_t sw, float scale,
                                 uint4* __restrict__ out, int B, int H, int W, int f16) {
  const int Wp = W + 8;
  const int64_t total = static_cast<int64_t>(B) * H * Wp;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xp = static_cast<int>(i % Wp);
    const int64_t t = i / Wp;
    const int y = static_cast<int>(t % H);
    const int64_t b = t / H;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    const int xi = xp - 3;
    if (xi >= 0 && xi < W) {
      const float* s = x + b * sb + y * sh + xi * sw;
      const float c0 = __ldg(s) * scale, c1 = __ldg(s + sc) * scale, c2 = __ldg(s + 2 * sc) * scale;
      o.x = pack16x2(c0, c1, f16 != 0);
      o.y = pack16x2(c2, 0.f, f16 != 0);
    }
    out[i] = o;
  }
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
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

// Tensor maps are a pure function of this key; the trunk calls with the same ~110 descriptions every step.
struct MapKey {
  const void* base;
  uint64_t dims[4];
  uint64_t strides[3];  // bytes, dims 1..3
  uint32_t box[4];
  uint32_t estr[4];
  int32_t rank, f16;
};
constexpr int kMapCacheSize = 256;
struct MapCache {
  MapKey key[kMapCacheSize];
  CUtensorMap map[kMapCacheSize];
  int used = 0, next = 0;
};
thread_local MapCache g_map_cache;

int make_map(CUtensorMap* tm, const MapKey& k) {
  MapCache& mc = g_map_cache;
  for (int i = 0; i < mc.used; ++i) {
    if (memcmp(&mc.key[i], &k, sizeof(MapKey)) == 0) {
      *tm = mc.map[i];
      return DUO_OK;
    }
  }
  PFN_encodeTiled fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DUO_ERR_CUDA;
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], estr[4];
  for (int i = 0; i < 4; ++i) {
    gdim[i] = k.dims[i];
    box[i] = k.box[i];
    estr[i] = k.estr[i];
  }
  for (int i = 0; i < 3; ++i) gstr[i] = k.strides[i];
  CUresult r = fn(tm, k.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  static_cast<cuuint32_t>(k.rank), const_cast<void*>(k.base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu %llu %llu %llu box %u %u %u %u estr %u %u %u %u)",
              (int)r, k.rank, (unsigned long long)k.dims[0], (unsigned long long)k.dims[1], (unsigned long long)k.dims[2],
              (unsigned long long)k.dims[3], k.box[0], k.box[1], k.box[2], k.box[3], k.estr[0], k.estr[1], k.estr[2], k.estr[3]);
    return DUO_ERR_CUDA;
  }
  const int slot = mc.used < kMapCacheSize ? mc.used++ : (mc.next = (mc.next + 1) % kMapCacheSize);
  mc.key[slot] = k;
  mc.map[slot] = *tm;
  return DUO_OK;
}

MapKey zero_key() {
  MapKey k;
  memset(&k, 0, sizeof(k));
  for (int i = 0; i < 4; ++i) {
    k.dims[i] = 1;
    k.box[i] = 1;
    k.estr[i] = 1;
  }
  return k;
}

int pow2_divisor_log2(int v, int cap_log2) {
  int l = 0;
  while (l < cap_log2 && (v & ((2 << l) - 1)) == 0) ++l;
  return l;
}

template <int BLOCK_N>
int launch_conv(const CUtensorMap& ti, const CUtensorMap& tw, const CUtensorMap& to, const ConvParams& p, cudaStream_t st) {
  using C = ConvCfg<BLOCK_N>;
  static uint64_t configured = 0;  // per device
  auto kfn = conv_tcgen05_kernel<BLOCK_N>;
  if (first_use_on_device(configured))
    DUO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(C::kSmemBytes)));
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;
  const int sms = device_sm_count();
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  kfn<<<grid, kThreads, C::kSmemBytes, st>>>(ti, tw, to, p);
  DUO_LAUNCH_CHECK("conv_tcgen05_kernel");
  return DUO_OK;
}

// Shared tail of duo_conv2d / duo_stem_conv7x7: tile box, output / weight maps, launch.
int run_conv(const CUtensorMap& tmap_in, ConvParams& p, const void* weight, int64_t k_total, void* out, int bw_log2,
             int bh_log2, cudaStream_t st) {
  const int bb = kBlockM >> (bw_log2 + bh_log2);
  p.bw_log2 = bw_log2;
  p.bh_log2 = bh_log2;
  p.tiles_w = p.Wo >> bw_log2;
  p.tiles_h = p.Ho >> bh_log2;
  const int64_t m_blocks = static_cast<int64_t>(p.tiles_w) * p.tiles_h * ((p.B + bb - 1) / bb);
  DUO_CHECK_ARG(m_blocks < (int64_t(1) << 30), "duo_conv2d: too many tiles");
  p.num_m_blocks = static_cast<int32_t>(m_blocks);
  const int sms = device_sm_count();
  int block_n = 64;
  if (p.Cout % 256 == 0 && m_blocks * (p.Cout / 256) >= 2 * sms) block_n = 256;
  else if (p.Cout % 128 == 0) block_n = 128;
  p.num_n_blocks = p.Cout / block_n;

  CUtensorMap tw, to;
  MapKey kw = zero_key();
  kw.base = weight;
  kw.rank = 2;
  kw.f16 = p.in_f16;
  kw.dims[0] = static_cast<uint64_t>(k_total);
  kw.dims[1] = static_cast<uint64_t>(p.Cout);
  kw.strides[0] = static_cast<uint64_t>(k_total) * 2;
  kw.box[0] = kBlockK;
  kw.box[1] = static_cast<uint32_t>(block_n);
  int rc = make_map(&tw, kw);
  if (rc != DUO_OK) return rc;
  // output: dims (Cout, Wo, Ho, B); box = one warp's slab of 32 pixels x 64 channels
  const int sw_log2 = bw_log2 < 5 ? bw_log2 : 5;
  const int sh_log2 = (bh_log2 < 5 - sw_log2) ? bh_log2 : 5 - sw_log2;
  const int sb_n = 32 >> (sw_log2 + sh_log2);
  MapKey ko = zero_key();
  ko.base = out;
  ko.rank = 4;
  ko.f16 = p.f16;
  ko.dims[0] = static_cast<uint64_t>(p.Cout);
  ko.dims[1] = static_cast<uint64_t>(p.Wo);
  ko.dims[2] = static_cast<uint64_t>(p.Ho);
  ko.dims[3] = static_cast<uint64_t>(p.B);
  ko.strides[0] = static_cast<uint64_t>(p.Cout) * 2;
  ko.strides[1] = ko.strides[0] * p.Wo;
  ko.strides[2] = ko.strides[1] * p.Ho;
  ko.box[0] = 64;
  ko.box[1] = 1u << sw_log2;
  ko.box[2] = 1u << sh_log2;
  ko.box[3] = static_cast<uint32_t>(sb_n);
  rc = make_map(&to, ko);
  if (rc != DUO_OK) return rc;
  switch (block_n) {
    case 256: return launch_conv<256>(tmap_in, tw, to, p, st);
    case 128: return launch_conv<128>(tmap_in, tw, to, p, st);
    default: return launch_conv<64>(tmap_in, tw, to, p, st);
  }
}

}  // namespace
}  // namespace duo

extern "C" int duo_conv2d(const duo_conv2d_args* a, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(a != nullptr, "duo_conv2d: args is NULL");
  DUO_CHECK_ARG(a->in && a->weight && a->out, "duo_conv2d: NULL operand");
  DUO_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0, "duo_conv2d: empty input B=%d H=%d W=%d", a->B, a->H, a->W);
  DUO_CHECK_ARG(a->Cin > 0 && a->Cin % 64 == 0 && a->Cout > 0 && a->Cout % 64 == 0,
                "duo_conv2d: channel counts must be multiples of 64 (Cin=%d Cout=%d)", a->Cin, a->Cout);
  DUO_CHECK_ARG(a->ksize == 1 || a->ksize == 3, "duo_conv2d: ksize=%d (1 or 3)", a->ksize);
  DUO_CHECK_ARG(a->stride == 1 || a->stride == 2, "duo_conv2d: stride=%d (1 or 2)", a->stride);
  DUO_CHECK_ARG(((reinterpret_cast<uintptr_t>(a->in) | reinterpret_cast<uintptr_t>(a->weight) |
                  reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->residual)) & 15) == 0,
                "duo_conv2d: tensors must be 16-byte aligned");
  DUO_CHECK_ARG(a->out != a->in && a->out != a->residual, "duo_conv2d: out must not alias in / residual");
  const int pad = a->ksize / 2;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.bias = a->bias;
  p.residual = a->residual;
  p.B = a->B;
  p.Ho = (a->H + 2 * pad - a->ksize) / a->stride + 1;
  p.Wo = (a->W + 2 * pad - a->ksize) / a->stride + 1;
  p.Cout = a->Cout;
  p.cin_blocks = a->Cin / 64;
  p.taps_x = p.taps_y = a->ksize;
  p.mul_w = p.mul_h = a->stride;
  p.off_w = p.off_h = -pad;
  p.relu = a->relu;
  p.f16 = a->out_fp16 ? 1 : 0;
  p.in_f16 = a->fp16 ? 1 : 0;
  p.idesc_mask = a->fp16 ? ~((1u << 7) | (1u << 10)) : ~0u;  // a_format / b_format: 1 = BF16, 0 = F16
  // tile box: powers of two dividing the output map, w first (at most 16 wide), then h, the rest along the batch
  const int bw_log2 = pow2_divisor_log2(p.Wo, 4);
  const int bh_log2 = pow2_divisor_log2(p.Ho, 7 - bw_log2 < 4 ? 7 - bw_log2 : 4);
  const int bb = kBlockM >> (bw_log2 + bh_log2);

  CUtensorMap ti;
  MapKey ki = zero_key();
  ki.base = a->in;
  ki.rank = 4;
  ki.f16 = p.in_f16;
  ki.dims[0] = static_cast<uint64_t>(a->Cin);
  ki.dims[1] = static_cast<uint64_t>(a->W);
  ki.dims[2] = static_cast<uint64_t>(a->H);
  ki.dims[3] = static_cast<uint64_t>(a->B);
  ki.strides[0] = static_cast<uint64_t>(a->Cin) * 2;
  ki.strides[1] = ki.strides[0] * a->W;
  ki.strides[2] = ki.strides[1] * a->H;
  ki.box[0] = kBlockK;
  ki.box[1] = static_cast<uint32_t>((1 << bw_log2) * a->stride);
  ki.box[2] = static_cast<uint32_t>((1 << bh_log2) * a->stride);
  ki.box[3] = static_cast<uint32_t>(bb);
  ki.estr[1] = ki.estr[2] = static_cast<uint32_t>(a->stride);
  int rc = make_map(&ti, ki);
  if (rc != DUO_OK) return rc;
  return run_conv(ti, p, a->weight, static_cast<int64_t>(a->ksize) * a->ksize * a->Cin, a->out, bw_log2, bh_log2,
                  reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int duo_stem_pack(const float* x, int64_t stride_b, int64_t stride_c, int64_t stride_h, int64_t stride_w,
                             float scale, void* out, int32_t fp16, int32_t B, int32_t H, int32_t W,
                             duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0, "duo_stem_pack: bad arguments");
  DUO_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "duo_stem_pack: out must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * H * (W + 8);
  const int threads = 256;
  const int64_t want = (total + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  const int blocks = static_cast<int>(want < cap ? want : cap);
  stem_pack_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, stride_b, stride_c, stride_h, stride_w, scale, reinterpret_cast<uint4*>(out), B, H, W, fp16);
  DUO_LAUNCH_CHECK("stem_pack_kernel");
  return DUO_OK;
}

extern "C" int duo_stem_conv7x7(const void* packed, const void* weight, const float* bias, void* out, int32_t B, int32_t H,
                                int32_t W, int32_t Cout, int32_t relu, int32_t fp16, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(packed && weight && out, "duo_stem_conv7x7: NULL operand");
  DUO_CHECK_ARG(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "duo_stem_conv7x7: even H, W expected (H=%d W=%d)", H, W);
  DUO_CHECK_ARG(Cout > 0 && Cout % 64 == 0, "duo_stem_conv7x7: Cout=%d must be a multiple of 64", Cout);
  DUO_CHECK_ARG(((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "duo_stem_conv7x7: tensors must be 16-byte aligned");
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias;
  p.B = B;
  p.Ho = H / 2;
  p.Wo = W / 2;
  p.Cout = Cout;
  p.cin_blocks = 1;
  p.taps_x = 1;
  p.taps_y = 7;
  p.mul_w = 1;  // the output-column stride (2 pixels) is the tensor map's dim-1 stride
  p.mul_h = 2;
  p.off_w = 0;  // the 3 pad pixels are physically present in the packed tensor
  p.off_h = -3;
  p.relu = relu;
  p.f16 = p.in_f16 = fp16 ? 1 : 0;
  p.idesc_mask = fp16 ? ~((1u << 7) | (1u << 10)) : ~0u;
  const int bw_log2 = pow2_divisor_log2(p.Wo, 4);
  const int bh_log2 = pow2_divisor_log2(p.Ho, 7 - bw_log2 < 4 ? 7 - bw_log2 : 4);
  const int bb = kBlockM >> (bw_log2 + bh_log2);
  const int Wp = W + 8;
  CUtensorMap ti;
  MapKey ki = zero_key();
  ki.base = packed;
  ki.rank = 4;
  ki.f16 = p.f16;
  ki.dims[0] = 64;                               // 8 pixels x 8 channels of one filter row's window
  ki.dims[1] = static_cast<uint64_t>(p.Wo);      // output column: windows 2 pixels (32 B) apart, overlapping
  ki.dims[2] = static_cast<uint64_t>(H);
  ki.dims[3] = static_cast<uint64_t>(B);
  ki.strides[0] = 32;
  ki.strides[1] = static_cast<uint64_t>(Wp) * 16;
  ki.strides[2] = ki.strides[1] * H;
  ki.box[0] = 64;
  ki.box[1] = 1u << bw_log2;
  ki.box[2] = static_cast<uint32_t>(2 << bh_log2);
  ki.box[3] = static_cast<uint32_t>(bb);
  ki.estr[2] = 2;
  int rc = make_map(&ti, ki);
  if (rc != DUO_OK) return rc;
  return run_conv(ti, p, weight, 7 * 64, out, bw_log2, bh_log2, reinterpret_cast<cudaStream_t>(stream));
}
