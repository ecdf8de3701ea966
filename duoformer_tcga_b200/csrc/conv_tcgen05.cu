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
// ROW-PAIR tensor [B, H + 8, W + 8, 8] (3 pad pixels left / top, 5 right / bottom; element (R, X, 0..3) = channels of
// padded pixel (R, X), element (R, X, 4..7) = channels of padded pixel (R + 1, X); channel 3 zero) and the input tensor
// map describes OVERLAPPING windows — dim0 = 64 elements (8 pixels x 2 rows x 4 channels, 128 contiguous bytes), dim1 =
// output column with a 2-pixel (32 B) stride, dim2 = padded image row (element stride 2), dim3 = batch — so TWO filter
// rows are one 64-wide K block: K = 4 * 64 for 7 x 7 x 3 = 147 real taps (the 8th filter row / column and the 4th
// channel meet zero weights).  (A 5-D map with the row pair as its own dimension and a 64-byte inner box was tried
// first: cuTensorMapEncodeTiled accepts it, the loads return garbage.)
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
constexpr int kEpiWarps = 8;                    // two groups of four: group e owns accumulator buffer e (every other tile)
constexpr int kThreads = 32 * (kEpiWarps + 2);  // + TMA producer warp + MMA issuer warp
// Per epilogue warp a ring of kSlots staging slots (32 pixel rows x 64 channels, 4 KB, SWIZZLE_128B).  Without a residual
// a slot is filled and TMA-stored; with one, the residual chunk is TMA-LOADED into the slot kSlots - 1 chunks ahead
// (also across tile boundaries, so the reads overlap the MMAs of the tile), updated in place and stored — the residual
// never travels through per-thread global loads (first version: 64 B per thread and row, 2.9 TB/s).
constexpr uint32_t kSlotBytes = 32 * 128;

template <int BLOCK_N>
struct ConvCfg {
  // 128 x 256 tiles (Cout % 256 == 0 and at least two waves of them): a third less L2 -> SM operand traffic per FLOP than
  // 128 x 128 — the 128-wide tiles of layers 3 - 4 sit at ~830 TFLOP/s = ~13 TB/s of operand traffic, the L2 limit; the
  // 48 KB stages leave room for three of them and two staging slots per warp.
  static constexpr int kStages = BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 4 : 5);
  static constexpr int kSlots = BLOCK_N == 256 ? 2 : 3;
  static constexpr uint32_t kStagingBytesPerWarp = kSlots * kSlotBytes;
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static constexpr uint32_t kStagingBytes = kEpiWarps * kStagingBytesPerWarp;
  static constexpr uint32_t kBarrierBytes = (2 * kStages + 4 + kEpiWarps * kSlots) * 8 + 8;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
};

// x / d for x * d < 2^40 (tile counts): one 32 x 32 -> 64 bit multiply and a shift instead of an integer division
struct FastDiv {
  uint64_t m;  // ceil(2^40 / d)
  uint32_t d;
  __host__ void init(uint32_t div) {
    d = div;
    m = ((uint64_t(1) << 40) + div - 1) / div;
  }
  // floor(x / d) = (x * m) >> 40, taken as the high half of (x << 24) * m
  __device__ __forceinline__ uint32_t div(uint32_t x) const {
    return static_cast<uint32_t>(__umul64hi(static_cast<uint64_t>(x) << 24, m));
  }
};

struct ConvParams {
  const float* bias;
  int32_t B, Ho, Wo, Cout;
  int32_t cin_blocks;        // K blocks per filter tap
  int32_t taps_x, taps_y;    // filter window walked by the K loop
  int32_t tap_step_h;        // input rows per filter-row step of the K loop (1; 2 for the stem's row-pair K blocks)
  int32_t mul_w, mul_h;      // input coordinate of tap (kx, ky) for output (w, h): w * mul_w + kx + off_w
  int32_t off_w, off_h;
  int32_t cin2_blocks, mul2; // fused projection shortcut: K blocks of a 1x1 / stride mul2 convolution of tmap_in2, appended
  int32_t bw_log2, bh_log2;  // M tile = 2^bw_log2 x 2^bh_log2 x (128 >> (bw_log2 + bh_log2)) output pixels (w, h, batch)
  int32_t num_m_blocks, num_n_blocks;
  FastDiv div_n, div_w, div_h;  // by num_n_blocks, tiles_w, tiles_h
  int32_t relu;
  uint32_t idesc_mask;
};

struct TileCoords {
  int n0, w0, h0, b0;
};
template <int BLOCK_N>
__device__ __forceinline__ TileCoords tile_coords(const ConvParams& p, uint32_t tile) {
  const uint32_t m_blk = p.div_n.div(tile);
  const uint32_t n_blk = tile - m_blk * p.div_n.d;
  const uint32_t t1 = p.div_w.div(m_blk);
  const uint32_t tw = m_blk - t1 * p.div_w.d;
  const uint32_t tb = p.div_h.div(t1);
  const uint32_t th = t1 - tb * p.div_h.d;
  TileCoords c;
  c.n0 = static_cast<int>(n_blk) * BLOCK_N;
  c.w0 = static_cast<int>(tw) << p.bw_log2;
  c.h0 = static_cast<int>(th) << p.bh_log2;
  c.b0 = static_cast<int>(tb) * (kBlockM >> (p.bw_log2 + p.bh_log2));
  return c;
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

template <bool F16>
__device__ __forceinline__ uint64_t unpack16x2(uint32_t u) {  // two 16-bit values -> packed fp32 pair
  if constexpr (F16) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&u));
    return pack2(t.x, t.y);
  } else {
    return pack2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
}
// packed fp32 pair -> two 16-bit values, clamped from below at `lo` (0 for ReLU, -inf otherwise: one HMNMX2 either way)
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2_max(uint64_t v, uint32_t lo) {
  float a, b;
  unpack2(v, a, b);
  if constexpr (F16) {
    const __half2 h = __hmax2(__floats2half2_rn(a, b), *reinterpret_cast<const __half2*>(&lo));
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(a, b), *reinterpret_cast<const __nv_bfloat162*>(&lo));
    return *reinterpret_cast<const uint32_t*>(&h);
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

template <int BLOCK_N, bool F16OUT, bool HAS_RES>
__global__ void __launch_bounds__(kThreads, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                    const __grid_constant__ CUtensorMap tmap_in2, const ConvParams p) {
  using C = ConvCfg<BLOCK_N>;
  constexpr int kStages = C::kStages;
  constexpr int kSlots = C::kSlots;
  constexpr uint32_t kStagingBytesPerWarp = C::kStagingBytesPerWarp;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + kStages * C::kStageBytes;
  const uint32_t bar_base = staging_base + C::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);
  auto res_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 5 + i); };  // "residual chunk loaded", one per slot
  uint32_t* tmem_ptr_generic = reinterpret_cast<uint32_t*>(smem_raw + (tmem_ptr_smem - ptx::smem_u32(smem_raw)));

  constexpr int kTmaWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;  // highest warp ids: never starved of issue slots
  const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == kTmaWarp && lane == 0) {
    ptx::prefetch_tmap(&tmap_in);
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_out);
    if constexpr (HAS_RES) ptx::prefetch_tmap(&tmap_res);
    if (p.cin2_blocks > 0) ptx::prefetch_tmap(&tmap_in2);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar(a), 1);
      ptx::mbar_init(tmem_empty_bar(a), 4);  // one arrival per epilogue warp of the buffer's group
    }
    for (int i = 0; i < kEpiWarps * kSlots; ++i) ptx::mbar_init(res_bar(i), 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == kMmaWarp) ptx::tmem_alloc<C::kTmemCols>(tmem_ptr_smem);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  const int num_k_blocks = p.taps_x * p.taps_y * p.cin_blocks + p.cin2_blocks;
  const uint32_t num_tiles = static_cast<uint32_t>(p.num_m_blocks) * static_cast<uint32_t>(p.num_n_blocks);

  if (warp_idx == kTmaWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next_stage = [&](uint32_t& sa, uint32_t& sb) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        sa = smem_base + stage * C::kStageBytes;
        sb = sa + C::kABytes;
        ptx::mbar_arrive_expect_tx(full_bar(stage), C::kStageBytes);
      };
      auto advance = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      };
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoords tc = tile_coords<BLOCK_N>(p, tile);
        const int cw = tc.w0 * p.mul_w + p.off_w;
        const int ch = tc.h0 * p.mul_h + p.off_h;
        int kb = 0;
        for (int ky = 0; ky < p.taps_y; ++ky) {
          for (int kx = 0; kx < p.taps_x; ++kx) {
            for (int cb = 0; cb < p.cin_blocks; ++cb, ++kb) {
              uint32_t sa, sb;
              next_stage(sa, sb);
              tma_load_4d(sa, &tmap_in, full_bar(stage), cb * kBlockK, cw + kx, ch + p.tap_step_h * ky, tc.b0);
              ptx::tma_load_2d(sb, &tmap_w, full_bar(stage), kb * kBlockK, tc.n0);
              advance();
            }
          }
        }
        for (int cb = 0; cb < p.cin2_blocks; ++cb, ++kb) {  // fused 1x1 projection shortcut of the block input
          uint32_t sa, sb;
          next_stage(sa, sb);
          tma_load_4d(sa, &tmap_in2, full_bar(stage), cb * kBlockK, tc.w0 * p.mul2, tc.h0 * p.mul2, tc.b0);
          ptx::tma_load_2d(sb, &tmap_w, full_bar(stage), kb * kBlockK, tc.n0);
          advance();
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
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
    // ===================== epilogue warps: group e = warp / 4 takes this CTA's tiles e, e + 2, ... =====================
    const int quarter = warp_idx & 3;  // TMEM lane quarter
    const int grp = warp_idx >> 2;     // accumulator buffer
    const uint32_t ring = staging_base + static_cast<uint32_t>(warp_idx) * kStagingBytesPerWarp;
    const uint32_t my_row_off = static_cast<uint32_t>(lane) * 128u;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    uint32_t acc_phase = 0;
    const uint32_t relu_lo = p.relu ? 0u : (F16OUT ? 0xFC00FC00u : 0xFF80FF80u);  // 0 or -inf, packed pair
    const int bwm = (1 << p.bw_log2) - 1, bhm = (1 << p.bh_log2) - 1;
    const int r0 = quarter * 32;  // first pixel of this warp's slab (a sub-box of the tile box)
    const int dw = r0 & bwm, dh = (r0 >> p.bw_log2) & bhm, db = r0 >> (p.bw_log2 + p.bh_log2);
    constexpr int kChunks = BLOCK_N / 64;
    const uint32_t tile_first = blockIdx.x + static_cast<uint32_t>(grp) * gridDim.x;
    const uint32_t tile_step = 2u * gridDim.x;
    // residual prefetch cursor (lane 0): next chunk to load
    uint32_t ld_tile = tile_first, ld_g = 0;
    int ld_ci = 0;
    TileCoords ld_tc = {0, 0, 0, 0};
    auto issue_load = [&]() {  // lane 0 only
      if (ld_tile >= num_tiles) return;
      if (ld_ci == 0) ld_tc = tile_coords<BLOCK_N>(p, ld_tile);
      const uint32_t s = ld_g % kSlots;
      const uint32_t bar = res_bar(warp_idx * kSlots + static_cast<int>(s));
      ptx::mbar_arrive_expect_tx(bar, kSlotBytes);
      tma_load_4d(ring + s * kSlotBytes, &tmap_res, bar, ld_tc.n0 + ld_ci * 64, ld_tc.w0 + dw, ld_tc.h0 + dh, ld_tc.b0 + db);
      ++ld_g;
      if (++ld_ci == kChunks) {
        ld_ci = 0;
        ld_tile += tile_step;
      }
    };
    if constexpr (HAS_RES) {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i + 1 < kSlots; ++i) issue_load();
      }
    }
    uint32_t g = 0;
    for (uint32_t tile = tile_first; tile < num_tiles; tile += tile_step) {
      const TileCoords tc = tile_coords<BLOCK_N>(p, tile);
      ptx::mbar_wait(tmem_full_bar(grp), acc_phase);
      ptx::tc_fence_after();
      acc_phase ^= 1u;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(grp * BLOCK_N);
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci, ++g) {
        const int c = ci * 64;
        const uint32_t s = g % kSlots;
        const uint32_t buf = ring + s * kSlotBytes + my_row_off;
        __syncwarp();  // lane 0 has seen the slot's previous store read out (wait_group.read below)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          float4 bia[8];
          ptx::tmem_ld_32x32(taddr + static_cast<uint32_t>(c + 32 * h), v);
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + tc.n0 + c + 32 * h);
#pragma unroll
            for (int j = 0; j < 8; ++j) bia[j] = __ldg(b4 + j);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) bia[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if constexpr (HAS_RES) {
            if (h == 0) ptx::mbar_wait(res_bar(warp_idx * kSlots + static_cast<int>(s)), (g / kSlots) & 1u);
          }
          ptx::tmem_ld_wait();
          if (h == 1 && ci == kChunks - 1) {  // accumulator fully read: hand the TMEM buffer back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(grp));
          }
          uint64_t t[16];  // packed fp32 pairs (FADD2)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            t[2 * j + 0] = add2(pack2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack2(bia[j].x, bia[j].y));
            t[2 * j + 1] = add2(pack2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack2(bia[j].z, bia[j].w));
          }
          if constexpr (HAS_RES) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // 16-byte chunk (4h + j) of this pixel row of the residual, XOR-swizzled
              uint32_t u[4];
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3])
                           : "r"(buf + ((static_cast<uint32_t>(4 * h + j) ^ sw) << 4))
                           : "memory");
#pragma unroll
              for (int i = 0; i < 4; ++i) t[4 * j + i] = add2(t[4 * j + i], unpack16x2<F16OUT>(u[i]));
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)  // in place: the same 16-byte chunks this thread just read
            st_shared_v4(buf + ((static_cast<uint32_t>(4 * h + j) ^ sw) << 4), pack16x2_max<F16OUT>(t[4 * j + 0], relu_lo),
                         pack16x2_max<F16OUT>(t[4 * j + 1], relu_lo), pack16x2_max<F16OUT>(t[4 * j + 2], relu_lo),
                         pack16x2_max<F16OUT>(t[4 * j + 3], relu_lo));
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmap_out, ring + s * kSlotBytes, tc.n0 + c, tc.w0 + dw, tc.h0 + dh, tc.b0 + db);
          ptx::tma_store_commit();
          ptx::tma_store_wait_read<1>();      // the previous chunk's store has left its slot ...
          if constexpr (HAS_RES) issue_load();  // ... which is the one chunk g + kSlots - 1 uses
        }
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

// Image -> zero-padded row-pair operand of the stem convolution ([B, H + 8, W + 8, 8]):
//   out[b, R, X, c] = scale * x[b, c, R - 3, X - 3],  out[b, R, X, 4 + c] = scale * x[b, c, R - 2, X - 3]  for c < 3,
// zero outside the image and for c = 3; any input strides (NCHW or channels-last), one 16-byte store per pixel.
__global__ void stem_pack_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sh, int64_t sw, float scale,
                                 uint4* __restrict__ out, int B, int H, int W, int f16) {
  const int Wp = W + 8, Hp = H + 8;
  const int64_t total = static_cast<int64_t>(B) * Hp * Wp;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xi = static_cast<int>(i % Wp) - 3;
    const int64_t t = i / Wp;
    const int yi = static_cast<int>(t % Hp) - 3;
    const int64_t b = t / Hp;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (xi >= 0 && xi < W) {
      const float* s = x + b * sb + xi * sw;
      if (yi >= 0 && yi < H) {
        const float* r = s + yi * sh;
        o.x = pack16x2(__ldg(r) * scale, __ldg(r + sc) * scale, f16 != 0);
        o.y = pack16x2(__ldg(r + 2 * sc) * scale, 0.f, f16 != 0);
      }
      if (yi + 1 >= 0 && yi + 1 < H) {
        const float* r = s + (yi + 1) * sh;
        o.z = pack16x2(__ldg(r) * scale, __ldg(r + sc) * scale, f16 != 0);
        o.w = pack16x2(__ldg(r + 2 * sc) * scale, 0.f, f16 != 0);
      }
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
  uint64_t dims[5];
  uint64_t strides[4];  // bytes, dims 1..4
  uint32_t box[5];
  uint32_t estr[5];
  int32_t rank, f16;
};
// Open-addressing hash table (linear probing, FNV-1a over the key bytes): a trunk pass looks up ~200 maps, a linear scan
// of the cache would put ~10^4 key comparisons on every forward's launch path.
constexpr int kMapCacheSize = 512;  // slots (power of two); the table is flushed when 3/4 full
struct MapCache {
  MapKey key[kMapCacheSize];
  CUtensorMap map[kMapCacheSize];
  bool used[kMapCacheSize];
  int count = 0;
  MapCache() { memset(used, 0, sizeof(used)); }
};
thread_local MapCache g_map_cache;

uint32_t key_hash(const MapKey& k) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof(MapKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
  return static_cast<uint32_t>(h ^ (h >> 32));
}

int make_map(CUtensorMap* tm, const MapKey& k) {
  MapCache& mc = g_map_cache;
  uint32_t slot = key_hash(k) & (kMapCacheSize - 1);
  while (mc.used[slot]) {
    if (memcmp(&mc.key[slot], &k, sizeof(MapKey)) == 0) {
      *tm = mc.map[slot];
      return DUO_OK;
    }
    slot = (slot + 1) & (kMapCacheSize - 1);
  }
  PFN_encodeTiled fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DUO_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t box[5], estr[5];
  for (int i = 0; i < 5; ++i) {
    gdim[i] = k.dims[i];
    box[i] = k.box[i];
    estr[i] = k.estr[i];
  }
  for (int i = 0; i < 4; ++i) gstr[i] = k.strides[i];
  CUresult r = fn(tm, k.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  static_cast<cuuint32_t>(k.rank), const_cast<void*>(k.base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu %llu %llu %llu %llu box %u %u %u %u %u)",
              (int)r, k.rank, (unsigned long long)k.dims[0], (unsigned long long)k.dims[1], (unsigned long long)k.dims[2],
              (unsigned long long)k.dims[3], (unsigned long long)k.dims[4], k.box[0], k.box[1], k.box[2], k.box[3], k.box[4]);
    return DUO_ERR_CUDA;
  }
  if (mc.count >= kMapCacheSize * 3 / 4) {  // activation buffers moved too often: start over
    memset(mc.used, 0, sizeof(mc.used));
    mc.count = 0;
    slot = key_hash(k) & (kMapCacheSize - 1);
  }
  mc.used[slot] = true;
  mc.key[slot] = k;
  mc.map[slot] = *tm;
  ++mc.count;
  return DUO_OK;
}

MapKey zero_key() {
  MapKey k;
  memset(&k, 0, sizeof(k));
  for (int i = 0; i < 5; ++i) {
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

template <int BLOCK_N, bool F16OUT, bool HAS_RES>
int launch_conv(const CUtensorMap& ti, const CUtensorMap& tw, const CUtensorMap& to, const CUtensorMap& tr,
                const CUtensorMap& ti2, const ConvParams& p, cudaStream_t st) {
  using C = ConvCfg<BLOCK_N>;
  static uint64_t configured = 0;  // per device
  auto kfn = conv_tcgen05_kernel<BLOCK_N, F16OUT, HAS_RES>;
  if (first_use_on_device(configured))
    DUO_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(C::kSmemBytes)));
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * p.num_n_blocks;
  const int sms = device_sm_count();
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  kfn<<<grid, kThreads, C::kSmemBytes, st>>>(ti, tw, to, tr, ti2, p);
  DUO_LAUNCH_CHECK("conv_tcgen05_kernel");
  return DUO_OK;
}

template <int BLOCK_N>
int dispatch_conv(bool f16out, bool has_res, const CUtensorMap& ti, const CUtensorMap& tw, const CUtensorMap& to,
                  const CUtensorMap& tr, const CUtensorMap& ti2, const ConvParams& p, cudaStream_t st) {
  if (f16out) return has_res ? launch_conv<BLOCK_N, true, true>(ti, tw, to, tr, ti2, p, st)
                             : launch_conv<BLOCK_N, true, false>(ti, tw, to, tr, ti2, p, st);
  return has_res ? launch_conv<BLOCK_N, false, true>(ti, tw, to, tr, ti2, p, st)
                 : launch_conv<BLOCK_N, false, false>(ti, tw, to, tr, ti2, p, st);
}

// Shared tail of duo_conv2d / duo_stem_conv7x7: tile box, output / weight / residual maps, launch.
// `tmap_in2` is only read when p.cin2_blocks > 0.
int run_conv(const CUtensorMap& tmap_in, const CUtensorMap& tmap_in2, ConvParams& p, const void* weight, int64_t k_total,
             void* out, const void* residual, bool in_f16, bool out_f16, int bw_log2, int bh_log2, cudaStream_t st) {
  const int bb = kBlockM >> (bw_log2 + bh_log2);
  p.bw_log2 = bw_log2;
  p.bh_log2 = bh_log2;
  const int tiles_w = p.Wo >> bw_log2, tiles_h = p.Ho >> bh_log2;
  const int64_t m_blocks = static_cast<int64_t>(tiles_w) * tiles_h * ((p.B + bb - 1) / bb);
  p.num_m_blocks = static_cast<int32_t>(m_blocks);
  int block_n = p.Cout % 128 == 0 ? 128 : 64;
  // (with a residual the two staging slots per warp of the 256-wide variant prefetch too little: 0.241 -> 0.257 ms on
  // layer 2's conv3; those keep 128-wide tiles and three slots)
  if (residual == nullptr && p.Cout % 256 == 0 && m_blocks * (p.Cout / 256) >= 2 * static_cast<int64_t>(device_sm_count()))
    block_n = 256;
  p.num_n_blocks = p.Cout / block_n;
  const int64_t tiles = m_blocks * p.num_n_blocks;
  // FastDiv: x * d < 2^40 for every quotient taken (x <= tiles, d <= 2^15)
  DUO_CHECK_ARG(tiles < (int64_t(1) << 24) && p.num_n_blocks < (1 << 15) && tiles_w < (1 << 15) && tiles_h < (1 << 15),
                "duo_conv2d: too many tiles (%lld)", (long long)tiles);
  p.div_n.init(static_cast<uint32_t>(p.num_n_blocks));
  p.div_w.init(static_cast<uint32_t>(tiles_w));
  p.div_h.init(static_cast<uint32_t>(tiles_h));

  CUtensorMap tw, to;
  MapKey kw = zero_key();
  kw.base = weight;
  kw.rank = 2;
  kw.f16 = in_f16;
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
  ko.f16 = out_f16;
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
  CUtensorMap tr = to;  // residual: the output's geometry on another base
  if (residual != nullptr) {
    ko.base = residual;
    rc = make_map(&tr, ko);
    if (rc != DUO_OK) return rc;
  }
  if (block_n == 256) return dispatch_conv<256>(out_f16, residual != nullptr, tmap_in, tw, to, tr, tmap_in2, p, st);
  if (block_n == 128) return dispatch_conv<128>(out_f16, residual != nullptr, tmap_in, tw, to, tr, tmap_in2, p, st);
  return dispatch_conv<64>(out_f16, residual != nullptr, tmap_in, tw, to, tr, tmap_in2, p, st);
}

// Input map of an NHWC tensor read with convolution stride `stride` by tiles of (1 << bw_log2) x (1 << bh_log2) x bb pixels.
int input_map(CUtensorMap* tm, const void* base, int B, int H, int W, int C, int stride, int bw_log2, int bh_log2, bool f16) {
  MapKey ki = zero_key();
  ki.base = base;
  ki.rank = 4;
  ki.f16 = f16;
  ki.dims[0] = static_cast<uint64_t>(C);
  ki.dims[1] = static_cast<uint64_t>(W);
  ki.dims[2] = static_cast<uint64_t>(H);
  ki.dims[3] = static_cast<uint64_t>(B);
  ki.strides[0] = static_cast<uint64_t>(C) * 2;
  ki.strides[1] = ki.strides[0] * W;
  ki.strides[2] = ki.strides[1] * H;
  ki.box[0] = kBlockK;
  ki.box[1] = static_cast<uint32_t>((1 << bw_log2) * stride);
  ki.box[2] = static_cast<uint32_t>((1 << bh_log2) * stride);
  ki.box[3] = static_cast<uint32_t>(kBlockM >> (bw_log2 + bh_log2));
  ki.estr[1] = ki.estr[2] = static_cast<uint32_t>(stride);
  return make_map(tm, ki);
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
                  reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->residual) |
                  reinterpret_cast<uintptr_t>(a->in2)) & 15) == 0,
                "duo_conv2d: tensors must be 16-byte aligned");
  DUO_CHECK_ARG(a->out != a->in && a->out != a->residual && a->out != a->in2, "duo_conv2d: out must not alias an input");
  const int pad = a->ksize / 2;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.bias = a->bias;
  p.B = a->B;
  p.Ho = (a->H + 2 * pad - a->ksize) / a->stride + 1;
  p.Wo = (a->W + 2 * pad - a->ksize) / a->stride + 1;
  p.Cout = a->Cout;
  p.cin_blocks = a->Cin / 64;
  p.taps_x = p.taps_y = a->ksize;
  p.tap_step_h = 1;
  p.mul_w = p.mul_h = a->stride;
  p.off_w = p.off_h = -pad;
  p.relu = a->relu;
  p.idesc_mask = a->fp16 ? ~((1u << 7) | (1u << 10)) : ~0u;  // a_format / b_format: 1 = BF16, 0 = F16
  // tile box: powers of two dividing the output map, w first (at most 16 wide), then h, the rest along the batch
  const int bw_log2 = pow2_divisor_log2(p.Wo, 4);
  const int bh_log2 = pow2_divisor_log2(p.Ho, 7 - bw_log2 < 4 ? 7 - bw_log2 : 4);
  CUtensorMap ti, ti2;
  int rc = input_map(&ti, a->in, a->B, a->H, a->W, a->Cin, a->stride, bw_log2, bh_log2, a->fp16 != 0);
  if (rc != DUO_OK) return rc;
  ti2 = ti;
  int64_t k_total = static_cast<int64_t>(a->ksize) * a->ksize * a->Cin;
  if (a->in2 != nullptr) {  // fused 1x1 projection shortcut
    DUO_CHECK_ARG(a->Cin2 > 0 && a->Cin2 % 64 == 0 && (a->stride2 == 1 || a->stride2 == 2),
                  "duo_conv2d: in2 needs Cin2 %% 64 == 0 and stride2 1 or 2 (Cin2=%d stride2=%d)", a->Cin2, a->stride2);
    DUO_CHECK_ARG((a->H2 - 1) / a->stride2 + 1 == p.Ho && (a->W2 - 1) / a->stride2 + 1 == p.Wo,
                  "duo_conv2d: in2 [%d x %d] / stride %d does not produce the output map %d x %d", a->H2, a->W2, a->stride2, p.Ho, p.Wo);
    p.cin2_blocks = a->Cin2 / 64;
    p.mul2 = a->stride2;
    rc = input_map(&ti2, a->in2, a->B, a->H2, a->W2, a->Cin2, a->stride2, bw_log2, bh_log2, a->fp16 != 0);
    if (rc != DUO_OK) return rc;
    k_total += a->Cin2;
  }
  return run_conv(ti, ti2, p, a->weight, k_total, a->out, a->residual, a->fp16 != 0, a->out_fp16 != 0, bw_log2, bh_log2,
                  reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int duo_stem_pack(const float* x, int64_t stride_b, int64_t stride_c, int64_t stride_h, int64_t stride_w,
                             float scale, void* out, int32_t fp16, int32_t B, int32_t H, int32_t W,
                             duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0, "duo_stem_pack: bad arguments");
  DUO_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "duo_stem_pack: out must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * (H + 8) * (W + 8);
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
  p.taps_y = 4;  // K blocks: filter rows (0,1), (2,3), (4,5), (6, zero)
  p.tap_step_h = 2;
  p.mul_w = 1;   // the output-column stride (2 pixels) is the tensor map's dim-1 stride
  p.mul_h = 2;
  p.off_w = 0;   // the pad pixels / rows are physically present in the packed tensor
  p.off_h = 0;
  p.relu = relu;
  p.idesc_mask = fp16 ? ~((1u << 7) | (1u << 10)) : ~0u;
  const int bw_log2 = pow2_divisor_log2(p.Wo, 4);
  const int bh_log2 = pow2_divisor_log2(p.Ho, 7 - bw_log2 < 4 ? 7 - bw_log2 : 4);
  const int bb = kBlockM >> (bw_log2 + bh_log2);
  const uint64_t row_pitch = static_cast<uint64_t>(W + 8) * 16;  // bytes of one packed image row
  CUtensorMap ti;
  MapKey ki = zero_key();
  ki.base = packed;
  ki.rank = 4;
  ki.f16 = fp16 ? 1 : 0;
  ki.dims[0] = 64;                               // 8 pixels x (2 rows x 4 channels): 128 contiguous bytes
  ki.dims[1] = static_cast<uint64_t>(p.Wo);      // output column: windows 2 pixels (32 B) apart, overlapping
  ki.dims[2] = static_cast<uint64_t>(H + 8);     // padded image row
  ki.dims[3] = static_cast<uint64_t>(B);
  ki.strides[0] = 32;
  ki.strides[1] = row_pitch;
  ki.strides[2] = row_pitch * (H + 8);
  ki.box[0] = 64;
  ki.box[1] = 1u << bw_log2;
  ki.box[2] = static_cast<uint32_t>(2 << bh_log2);
  ki.box[3] = static_cast<uint32_t>(bb);
  ki.estr[2] = 2;
  int rc = make_map(&ti, ki);
  if (rc != DUO_OK) return rc;
  return run_conv(ti, ti, p, weight, 4 * 64, out, nullptr, fp16 != 0, fp16 != 0, bw_log2, bh_log2,
                  reinterpret_cast<cudaStream_t>(stream));
}
