// Small HBM-bound kernels around the token tensor: scale-token row, patch-stage input
// assembly (CLS + s=0 tokens + pos_embed), classification head, fp32 -> bf16 / split convert.
#include "common.cuh"

namespace duo {
namespace {

// X[b,p,0,:] = tok[b,p,:] + pos0[:]   (float4 per thread)
__global__ void fill_scale_token_kernel(float* __restrict__ X, const float* __restrict__ tok,
                                        int64_t tsb, int64_t tsp, const float* __restrict__ pos0,
                                        int B, int P, int S, int D) {
  const int d4 = D >> 2;
  const int64_t total = static_cast<int64_t>(B) * P * d4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d4);
    const int64_t bp = i / d4;
    const int p = static_cast<int>(bp % P);
    const int64_t b = bp / P;
    const float4 t = __ldg(reinterpret_cast<const float4*>(tok + b * tsb + p * tsp) + c);
    const float4 q = __ldg(reinterpret_cast<const float4*>(pos0) + c);
    reinterpret_cast<float4*>(X + (bp * S) * D)[c] =
        make_float4(t.x + q.x, t.y + q.y, t.z + q.z, t.w + q.w);
  }
}

// Z[b,0,:] = cls + pos[0];  Z[b,1+p,:] = X[b,p,0,:] + pos[1+p]
template <int OUT_KIND>
__global__ void assemble_patch_tokens_kernel(const float* __restrict__ X,
                                             const float* __restrict__ cls,
                                             const float* __restrict__ pos, void* __restrict__ Z,
                                             int B, int P, int S, int D) {
  const int d4 = D >> 2;
  const int N = P + 1;
  const int64_t total = static_cast<int64_t>(B) * N * d4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d4);
    const int64_t bn = i / d4;
    const int n = static_cast<int>(bn % N);
    const int64_t b = bn / N;
    float4 v;
    if (n == 0)
      v = __ldg(reinterpret_cast<const float4*>(cls) + c);
    else
      v = __ldg(reinterpret_cast<const float4*>(X + ((b * P + (n - 1)) * S) * D) + c);
    const float4 q = __ldg(reinterpret_cast<const float4*>(pos + static_cast<int64_t>(n) * D) + c);
    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    if constexpr (OUT_KIND == DUO_ACT_BF16) {
      uint2 w;
      w.x = pack_bf16x2(v.x, v.y);
      w.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(Z) + bn * D)[c] = w;
    } else {
      uint2 h, l;
      pack_split2(v.x, v.y, h.x, l.x);
      pack_split2(v.z, v.w, h.y, l.y);
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(Z) + bn * (2 * D);
      reinterpret_cast<uint2*>(o)[c] = h;
      reinterpret_cast<uint2*>(o + D)[c] = l;
    }
  }
}

// out[r,:] = in[r,:] + pos[r % S, :]
__global__ void add_pos_kernel(const float* __restrict__ in, const float* __restrict__ pos,
                               float* __restrict__ out, int64_t rows, int S, int D) {
  const int d4 = D >> 2;
  const int64_t total = rows * d4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d4);
    const int64_t r = i / d4;
    const int s = static_cast<int>(r % S);
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + r * D) + c);
    const float4 q = __ldg(reinterpret_cast<const float4*>(pos + static_cast<int64_t>(s) * D) + c);
    reinterpret_cast<float4*>(out + r * D)[c] = make_float4(v.x + q.x, v.y + q.y, v.z + q.z, v.w + q.w);
  }
}

// One CTA (128 threads) per image: optional LayerNorm of the row, then ncls dot products.
__global__ void __launch_bounds__(128)
head_kernel(const float* __restrict__ in, int64_t row_stride, const float* __restrict__ ln_g,
            const float* __restrict__ ln_b, float eps, const float* __restrict__ W,
            const float* __restrict__ bias, float* __restrict__ logits, int D, int ncls) {
  extern __shared__ float z[];  // D floats
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = in + static_cast<int64_t>(b) * row_stride;
  for (int d = tid; d < D; d += 128) z[d] = row[d];
  __syncthreads();
  if (ln_g != nullptr) {
    float s = 0.f;
    for (int d = tid; d < D; d += 128) s += z[d];
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    const float mean = (red[0] + red[1] + red[2] + red[3]) / D;
    __syncthreads();
    float q = 0.f;
    for (int d = tid; d < D; d += 128) {
      const float t = z[d] - mean;
      q += t * t;
    }
    q = warp_sum(q);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    const float rstd = rsqrtf((red[0] + red[1] + red[2] + red[3]) / D + eps);
    for (int d = tid; d < D; d += 128) z[d] = (z[d] - mean) * rstd * ln_g[d] + ln_b[d];
    __syncthreads();
  }
  for (int c = warp; c < ncls; c += 4) {
    const float* w = W + static_cast<int64_t>(c) * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(z[d], __ldg(w + d), s);
    s = warp_sum(s);
    if (lane == 0) logits[static_cast<int64_t>(b) * ncls + c] = s + (bias ? bias[c] : 0.f);
  }
}

template <int OUT_KIND>
__global__ void convert_kernel(const float* __restrict__ in, int64_t ld, void* __restrict__ out,
                               int64_t rows, int cols) {
  const int c4n = cols >> 2;
  const int64_t total = rows * c4n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c4n);
    const int64_t r = i / c4n;
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + r * ld) + c);
    if constexpr (OUT_KIND == DUO_ACT_BF16) {
      uint2 w;
      w.x = pack_bf16x2(v.x, v.y);
      w.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + r * cols)[c] = w;
    } else {
      uint2 h, l;
      pack_split2(v.x, v.y, h.x, l.x);
      pack_split2(v.z, v.w, h.y, l.y);
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + r * (2 * static_cast<int64_t>(cols));
      reinterpret_cast<uint2*>(o)[c] = h;
      reinterpret_cast<uint2*>(o + cols)[c] = l;
    }
  }
}

inline unsigned elementwise_grid(int64_t total_threads) {
  const int64_t want = (total_threads + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  return static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace duo

extern "C" int duo_fill_scale_token(float* X, const float* tok, int64_t tok_stride_b,
                                    int64_t tok_stride_p, const float* pos0, int32_t B, int32_t P,
                                    int32_t S, int32_t D, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(X && tok && pos0, "duo_fill_scale_token: NULL pointer");
  DUO_CHECK_ARG(B > 0 && P > 0 && S > 0 && D > 0 && D % 4 == 0, "duo_fill_scale_token: bad dims");
  DUO_CHECK_ARG(tok_stride_b % 4 == 0 && tok_stride_p % 4 == 0,
                "duo_fill_scale_token: token strides must be multiples of 4");
  const int64_t total = static_cast<int64_t>(B) * P * (D / 4);
  fill_scale_token_kernel<<<elementwise_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      X, tok, tok_stride_b, tok_stride_p, pos0, B, P, S, D);
  DUO_LAUNCH_CHECK("fill_scale_token_kernel");
  return DUO_OK;
}

extern "C" int duo_add_pos(const float* in, const float* pos, float* out, int64_t rows, int32_t S,
                           int32_t D, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(in && pos && out, "duo_add_pos: NULL pointer");
  DUO_CHECK_ARG(rows > 0 && S > 0 && D > 0 && D % 4 == 0, "duo_add_pos: bad dims");
  const int64_t total = rows * (D / 4);
  add_pos_kernel<<<elementwise_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, pos, out, rows, S, D);
  DUO_LAUNCH_CHECK("add_pos_kernel");
  return DUO_OK;
}

extern "C" int duo_assemble_patch_tokens(const float* X, const float* cls, const float* pos,
                                         void* Z, int32_t out_kind, int32_t B, int32_t P,
                                         int32_t S, int32_t D, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(X && cls && pos && Z, "duo_assemble_patch_tokens: NULL pointer");
  DUO_CHECK_ARG(B > 0 && P > 0 && S > 0 && D > 0 && D % 4 == 0, "duo_assemble_patch_tokens: bad dims");
  DUO_CHECK_ARG(out_kind == DUO_ACT_BF16 || out_kind == DUO_ACT_SPLIT,
                "duo_assemble_patch_tokens: out_kind=%d", out_kind);
  const int64_t total = static_cast<int64_t>(B) * (P + 1) * (D / 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_kind == DUO_ACT_BF16)
    assemble_patch_tokens_kernel<DUO_ACT_BF16><<<elementwise_grid(total), 256, 0, st>>>(X, cls, pos, Z, B, P, S, D);
  else
    assemble_patch_tokens_kernel<DUO_ACT_SPLIT><<<elementwise_grid(total), 256, 0, st>>>(X, cls, pos, Z, B, P, S, D);
  DUO_LAUNCH_CHECK("assemble_patch_tokens_kernel");
  return DUO_OK;
}

extern "C" int duo_head(const float* in, int64_t row_stride, const float* ln_gamma,
                        const float* ln_beta, float eps, const float* W, const float* bias,
                        float* logits, int32_t B, int32_t D, int32_t num_classes,
                        duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(in && W && logits, "duo_head: NULL pointer");
  DUO_CHECK_ARG(B > 0 && D > 0 && num_classes > 0 && D <= 8192, "duo_head: bad dims");
  DUO_CHECK_ARG((ln_gamma == nullptr) == (ln_beta == nullptr), "duo_head: need both LN params");
  head_kernel<<<B, 128, D * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      in, row_stride, ln_gamma, ln_beta, eps, W, bias, logits, D, num_classes);
  DUO_LAUNCH_CHECK("head_kernel");
  return DUO_OK;
}

extern "C" int duo_convert(const float* in, int64_t ld, void* out, int32_t out_kind, int64_t rows,
                           int32_t cols, duo_stream_t stream) {
  using namespace duo;
  DUO_CHECK_ARG(in && out, "duo_convert: NULL pointer");
  DUO_CHECK_ARG(rows > 0 && cols > 0 && cols % 4 == 0 && ld % 4 == 0 && ld >= cols,
                "duo_convert: bad dims rows=%lld cols=%d ld=%lld", (long long)rows, cols, (long long)ld);
  DUO_CHECK_ARG(out_kind == DUO_ACT_BF16 || out_kind == DUO_ACT_SPLIT, "duo_convert: out_kind=%d", out_kind);
  const int64_t total = rows * (cols / 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_kind == DUO_ACT_BF16)
    convert_kernel<DUO_ACT_BF16><<<elementwise_grid(total), 256, 0, st>>>(in, ld, out, rows, cols);
  else
    convert_kernel<DUO_ACT_SPLIT><<<elementwise_grid(total), 256, 0, st>>>(in, ld, out, rows, cols);
  DUO_LAUNCH_CHECK("convert_kernel");
  return DUO_OK;
}
