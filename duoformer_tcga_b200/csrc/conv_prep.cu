// Data-movement kernels that let 3x3 convolutions and 2x2 max-pools of the channel-token branch
// (projection_head.py:152-268; call sites model.py:279-289, model_wo_extra_params.py:236-248) run on
// the tcgen05 GEMM: NHWC im2col (-> bf16 [B*Ho*Wo, 9*C], column order (ky, kx, c)) and NHWC
// max-pool / copy into a channel slice of the concatenated [B*Ho*Wo, C_total] tensor.
// Both are HBM-bound, 16-byte vectorised, grid-stride.
#include <cuda_fp16.h>

#include "common.cuh"

namespace duo {
namespace {

template <typename T>
struct Vec8;  // 8 consecutive channels -> 8 floats
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <>
struct Vec8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&f)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// out[(b, yo, xo), (ky, kx, c)] = in[b, yo*stride + ky - 1, xo*stride + kx - 1, c]  (zero padding 1)
template <typename T>
__global__ void im2col3x3_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H,
                                 int W, int C, int stride, int Ho, int Wo) {
  const int c8n = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * 9 * c8n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % c8n);
    int64_t r = i / c8n;
    const int tap = static_cast<int>(r % 9);
    r /= 9;  // output pixel index (b, yo, xo)
    const int xo = static_cast<int>(r % Wo);
    const int yo = static_cast<int>((r / Wo) % Ho);
    const int64_t b = r / (static_cast<int64_t>(Wo) * Ho);
    const int y = yo * stride + tap / 3 - 1;
    const int x = xo * stride + tap % 3 - 1;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < H && x >= 0 && x < W)
      Vec8<T>::load(in + ((b * H + y) * W + x) * C + c8 * 8, f);
    reinterpret_cast<uint4*>(out + r * (9 * static_cast<int64_t>(C)) + tap * C)[c8] = pack8(f);
  }
}

// pool == 2: out[(b, yo, xo), c] = max over the 2x2 window;  pool == 1: plain copy / conversion.
// `out` points at the first channel of the slice; rows are ld_out elements apart.
template <typename T>
__global__ void pool_to_slice_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H,
                                     int W, int C, int pool, int64_t ld_out) {
  const int Ho = H / pool, Wo = W / pool;
  const int c8n = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * c8n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % c8n);
    const int64_t r = i / c8n;
    const int xo = static_cast<int>(r % Wo);
    const int yo = static_cast<int>((r / Wo) % Ho);
    const int64_t b = r / (static_cast<int64_t>(Wo) * Ho);
    float m[8];
    Vec8<T>::load(in + ((b * H + yo * pool) * W + xo * pool) * C + c8 * 8, m);
    if (pool == 2) {
#pragma unroll
      for (int t = 1; t < 4; ++t) {
        float f[8];
        Vec8<T>::load(in + ((b * H + yo * 2 + (t >> 1)) * W + xo * 2 + (t & 1)) * C + c8 * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
      }
    }
    reinterpret_cast<uint4*>(out + r * ld_out)[c8] = pack8(m);
  }
}

// MaxPool2d(kernel 3, stride 2, padding 1) on NHWC, same 16-bit type in and out (max of representable
// values is exact).  One thread per (output pixel, 8 channels): the 3x3 window is 9 coalesced 16-byte loads.
template <typename T2>  // __half2 or __nv_bfloat162
__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W,
                                    int C, int Ho, int Wo) {
  const int c8n = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * c8n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % c8n);
    const int64_t r = i / c8n;
    const int xo = static_cast<int>(r % Wo);
    const int yo = static_cast<int>((r / Wo) % Ho);
    const int64_t b = r / (static_cast<int64_t>(Wo) * Ho);
    const int y0 = 2 * yo - 1, x0 = 2 * xo - 1;
    bool have = false;
    uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = y0 + dy;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int x = x0 + dx;
        if (x < 0 || x >= W) continue;
        const uint4 v = __ldg(in + ((b * H + y) * W + x) * c8n + c8);
        if (!have) {
          m = v;
          have = true;
        } else {
          T2* mm = reinterpret_cast<T2*>(&m);
          const T2* vv = reinterpret_cast<const T2*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j) mm[j] = __hmax2(mm[j], vv[j]);
        }
      }
    }
    out[r * c8n + c8] = m;
  }
}

inline unsigned grid_for(int64_t total) {
  const int64_t want = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  return static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace duo

extern "C" int duo_im2col3x3(const void* in, int32_t in_kind, void* out, int32_t B, int32_t H, int32_t W,
                             int32_t C, int32_t stride, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(in && out, "duo_im2col3x3: NULL pointer");
  DUO_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
                "duo_im2col3x3: bad dims B=%d H=%d W=%d C=%d stride=%d", B, H, W, C, stride);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;  // kernel 3, padding 1
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * 9 * (C / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  switch (in_kind) {
    case DUO_ACT_BF16:
      im2col3x3_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), o, B, H, W, C, stride, Ho, Wo);
      break;
    case DUO_ACT_F16:
      im2col3x3_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const __half*>(in), o, B, H, W, C, stride, Ho, Wo);
      break;
    case DUO_ACT_F32:
      im2col3x3_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const float*>(in), o, B, H, W, C, stride, Ho, Wo);
      break;
    default: set_error("duo_im2col3x3: in_kind=%d", in_kind); return DUO_ERR_INVALID;
  }
  DUO_LAUNCH_CHECK("im2col3x3_kernel");
  return DUO_OK;
}

extern "C" int duo_maxpool3x3s2(const void* in, int32_t kind, void* out, int32_t B, int32_t H, int32_t W,
                                int32_t C, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(in && out, "duo_maxpool3x3s2: NULL pointer");
  DUO_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "duo_maxpool3x3s2: bad dims B=%d H=%d W=%d C=%d", B, H, W, C);
  DUO_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "duo_maxpool3x3s2: pointers must be 16-byte aligned");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;  // kernel 3, stride 2, padding 1
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * (C / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint4* i4 = reinterpret_cast<const uint4*>(in);
  uint4* o4 = reinterpret_cast<uint4*>(out);
  switch (kind) {
    case DUO_ACT_BF16: maxpool3x3s2_kernel<__nv_bfloat162><<<grid_for(total), 256, 0, st>>>(i4, o4, B, H, W, C, Ho, Wo); break;
    case DUO_ACT_F16: maxpool3x3s2_kernel<__half2><<<grid_for(total), 256, 0, st>>>(i4, o4, B, H, W, C, Ho, Wo); break;
    default: set_error("duo_maxpool3x3s2: kind=%d (bf16 or f16 only)", kind); return DUO_ERR_INVALID;
  }
  DUO_LAUNCH_CHECK("maxpool3x3s2_kernel");
  return DUO_OK;
}

extern "C" int duo_pool_to_slice(const void* in, int32_t in_kind, void* out, int64_t ld_out, int32_t B,
                                 int32_t H, int32_t W, int32_t C, int32_t pool, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(in && out, "duo_pool_to_slice: NULL pointer");
  DUO_CHECK_ARG(B > 0 && C > 0 && C % 8 == 0 && (pool == 1 || pool == 2) && H % pool == 0 && W % pool == 0 &&
                    ld_out >= C && ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "duo_pool_to_slice: bad dims B=%d H=%d W=%d C=%d pool=%d ld_out=%lld", B, H, W, C, pool, (long long)ld_out);
  const int64_t total = static_cast<int64_t>(B) * (H / pool) * (W / pool) * (C / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  switch (in_kind) {
    case DUO_ACT_BF16:
      pool_to_slice_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), o, B, H, W, C, pool, ld_out);
      break;
    case DUO_ACT_F16:
      pool_to_slice_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const __half*>(in), o, B, H, W, C, pool, ld_out);
      break;
    case DUO_ACT_F32:
      pool_to_slice_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const float*>(in), o, B, H, W, C, pool, ld_out);
      break;
    default: set_error("duo_pool_to_slice: in_kind=%d", in_kind); return DUO_ERR_INVALID;
  }
  DUO_LAUNCH_CHECK("pool_to_slice_kernel");
  return DUO_OK;
}
