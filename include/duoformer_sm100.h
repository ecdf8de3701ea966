/*
 * duoformer_sm100.h — C ABI of libduoformer_sm100.so
 *
 * B200 (sm_100a) kernels for the DuoFormer multi-scale transformer forward path.
 * The reference (AliSerwat/duoformer_TCGA) is pure PyTorch and has no FFI of its own; the
 * boundary it exposes is the Python module API of models/model.py and
 * models/model_wo_extra_params.py.  This C ABI sits directly beneath that API: every entry
 * point below replaces one group of ATen calls of the reference forward, cited per function
 * as <reference file>:<lines>.  The Python classes in duoformer_tcga_b200/ keep the
 * reference's constructor signatures / state_dict schema and call these functions through
 * ctypes (see INTEGRATION.md for the binding).
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (PyTorch caching allocator).  The
 *     library never allocates, frees or retains device memory.
 *   - Every call only ENQUEUES work on `stream` (a cudaStream_t); no host synchronisation, so
 *     calls are CUDA-graph capturable.
 *   - Return value: 0 = ok, negative = error (DUO_ERR_*); duo_last_error() returns a
 *     thread-local message.  No C++ exception crosses the ABI.
 *   - "bf16" = __nv_bfloat16 storage.  "split bf16" = a [rows, 2*cols] bf16 matrix holding
 *     hi = bf16(x) in columns [0, cols) and lo = bf16(x - hi) in [cols, 2*cols): the operand
 *     format of the 3-pass (hi*hi + hi*lo + lo*hi) tensor-core GEMM used for the fp32-accuracy
 *     mode (north-star tolerance 1e-3).
 *   - There is NO CPU fallback.
 */
#ifndef DUOFORMER_SM100_H_
#define DUOFORMER_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* duo_stream_t; /* cudaStream_t */

enum {
  DUO_OK = 0,
  DUO_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  DUO_ERR_CUDA = -2,        /* CUDA runtime / driver error (message has the code) */
  DUO_ERR_UNSUPPORTED = -3  /* device is not sm_100 */
};

/* GEMM epilogues (duo_gemm_args.epilogue) */
enum {
  DUO_EPI_BF16 = 0,            /* out bf16 [M,N]      = acc + bias                         */
  DUO_EPI_GELU_BF16 = 1,       /* out bf16 [M,N]      = gelu_erf(acc + bias)               */
  DUO_EPI_RESIDUAL_F32 = 2,    /* out f32  [M,N]     += gamma[n] * (acc + bias)  (in place) */
  DUO_EPI_SCATTER_F32 = 3,     /* out f32  [*,N]: row r -> dest(r); = acc + bias + pos     */
  DUO_EPI_F32 = 4,             /* out f32  [M,N]      = acc + bias                         */
  DUO_EPI_SPLIT_BF16 = 5,      /* out split bf16 [M,2N] of (acc + bias)                    */
  DUO_EPI_GELU_SPLIT_BF16 = 6  /* out split bf16 [M,2N] of gelu_erf(acc + bias)            */
};

/* Element kinds of activation tensors */
enum {
  DUO_ACT_BF16 = 0,  /* bf16 [rows, cols]        */
  DUO_ACT_SPLIT = 1, /* split bf16 [rows, 2*cols] */
  DUO_ACT_F32 = 2,   /* float [rows, cols]       */
  DUO_ACT_F16 = 3    /* __half [rows, cols] (input of the im2col / pool kernels: the fp16 cuDNN trunk) */
};

const char* duo_last_error(void);
int duo_abi_version(void);
/* Number of kernels launched by this library in the calling thread since the last reset
 * (bench.py's "gpu_launches"). */
int64_t duo_launch_count(void);
void duo_launch_count_reset(void);

/*
 * Dense contraction  C[M,N] = A[M,K] * W[N,K]^T  (+ epilogue), bf16 operands, fp32 accumulate
 * in TMEM, TMA-fed tcgen05.mma.  Replaces every nn.Linear / 1x1 Conv2d on the path:
 *   qkv / proj            scale_attention.py:31,42 ; multiscale_attn.py:144-146,151,164
 *   fc1 / GELU / fc2      timm Mlp, instantiated scale_attention.py:79-84
 *   1x1 conv projection   projection_head.py:134-149 (+ gather/cat/permute
 *                         model_wo_extra_params.py:252-299 via DUO_EPI_SCATTER_F32)
 *   residual / LayerScale scale_attention.py:91-92 ; multiscale_attn.py:282-285
 * Requirements: N % 128 == 0, K % 64 == 0, A/W 16-byte aligned, lda/ldw multiples of 8.
 * split3 == 1: A is split bf16 [M,2K], W is split bf16 [N,2K]; computes Ah*Wh + Ah*Wl + Al*Wh.
 * split3 == 2: A is plain bf16 [M,K] (exact), W is split bf16 [N,2K]; computes A*Wh + A*Wl.
 */
typedef struct duo_gemm_args {
  const void* A;       /* bf16 [M, K] (or [M, 2K] when split3) row-major, leading dim lda */
  const void* W;       /* bf16 [N, K] (or [N, 2K] when split3) row-major, leading dim ldw */
  const float* bias;   /* [N] or NULL */
  void* out;           /* see epilogue; leading dim ldo (elements of the out type)        */
  const float* gamma;  /* RESIDUAL: LayerScale [N] or NULL (== 1)                          */
  const int32_t* row_map; /* SCATTER: [rows_per_group] dest row inside the group           */
  const float* pos;    /* SCATTER: [pos_period, N] added by (dest_row % pos_period) or NULL */
  int64_t M;
  int64_t lda, ldw, ldo;
  int32_t N, K;
  int32_t split3;
  int32_t epilogue;
  int32_t rows_per_group;      /* SCATTER: source rows per image (h*w of the stage)        */
  int32_t dest_rows_per_group; /* SCATTER: token rows per image (P*S)                      */
  int32_t pos_period;          /* SCATTER: S                                               */
  float ln_eps;                /* forwarded LayerNorm epsilon (consumer side)              */
  int32_t relu;                /* BF16 / F32 epilogues: out = max(acc + bias, 0) (conv + BN + ReLU of   */
  int32_t fp16_operands;       /* the channel-token branch, projection_head.py:242-254).               */
                               /* fp16_operands = 1: A and W hold IEEE fp16 instead of bf16 (the cuDNN */
                               /* trunk maps are fed as they are; plain mode only, split3 == 0; one    */
                               /* tcgen05.mma takes A and B of the same 16-bit format)                 */
  /*
   * LayerNorm statistics forwarding — the `x = x + f(x); y = Linear(norm(x))` pairs of
   * scale_attention.py:91-92 without a LayerNorm pass over the residual stream:
   *
   * producer (RESIDUAL_F32, bf16 operands, N % 256 == 0, both pointers set): the epilogue loads the fp32
   *   rows of `out` itself (TMA), adds gamma * (acc + bias), stores them back and ALSO writes
   *     xb_out    bf16 [M, N] dense: the updated, un-normalised rows (A operand of the next GEMM), and
   *     stats_out float [M, N / 256, 2]: (mean, sum of squared deviations) of every 256-column part of
   *               the updated fp32 row.
   * consumer (BF16 / GELU_BF16, bf16 operands, K % 256 == 0, K <= 1024, ln_stats set): A is such an un-normalised
   *   copy, W holds W * diag(ln_weight) with every row CENTRED (sum_k W[n, k] = 0: the row mean of x then cancels
   *   inside the product, x W^T = (x - mean) W^T) and bias holds W ln_bias + b; the epilogue merges the row's K / 256
   *   partial statistics (ln_stats, same layout as stats_out) into rstd and computes
   *     out = rstd * acc + bias[n],
   *   which equals Linear(LayerNorm(x)) up to operand rounding.
   */
  void* xb_out;
  float* stats_out;
  const float* ln_stats;
  /* producer, optional: statistics (same layout) of the rows BEFORE this update — the previous forwarding GEMM's      */
  /* stats_out or duo_layernorm's.  xb_out then holds bf16(x - m), m = the row mean according to these statistics: the */
  /* consumer's row-centred weights make any per-row constant drop out of its product, and rounding x - m instead of  */
  /* x keeps the bf16 error relative to the row's spread when |mean| >> spread.  Must not alias stats_out.            */
  const float* shift_stats;
} duo_gemm_args;
int duo_gemm(const duo_gemm_args* args, duo_stream_t stream);

/*
 * LayerNorm over the last dim (fp32 statistics, two-pass), fp32 in -> bf16 / split bf16 out.
 * Replaces nn.LayerNorm(eps=1e-6): scale_attention.py:65,78,91-92; multiscale_attn.py:282-285.
 * dim % 128 == 0, dim <= 1024.  ldx = input row stride in elements (>= dim; rows of a strided view,
 * e.g. the s = 0 token of every patch, can be normalised without a gather); out is dense.
 * stats_out (optional, dim % 256 == 0): float [rows, dim / 256, 2], (mean, sum of squared deviations) of every
 * 256-column part of the row — the layout of duo_gemm's statistics forwarding (used as its shift_stats).
 */
int duo_layernorm(const float* x, const float* gamma, const float* beta, void* out,
                  int32_t out_kind, int64_t rows, int32_t dim, int64_t ldx, float eps,
                  float* stats_out, duo_stream_t stream);

/*
 * Grouped multi-head attention over S consecutive rows of a fused qkv matrix:
 *   for every group g (S rows) and head h: out = softmax(q k^T * scale) v, heads merged.
 * qkv row layout [3][H][64] (which*D + h*64 + d), exactly the reference's
 * reshape(..., 3, H, dh): scale_attention.py:30-41 (scale attention, group = one patch,
 * S = 6/22/86), scale_attention.py:195-207 / multiscale_attn.py:205-216 (patch attention,
 * group = one image, S = P+1).  head_dim must be 64.
 * in_kind: DUO_ACT_BF16 or DUO_ACT_F32; out_kind: DUO_ACT_BF16 / DUO_ACT_SPLIT / DUO_ACT_F32.
 * in_kind DUO_ACT_SPLIT (qkv rows = [hi: q k v | lo: q k v], the SPLIT epilogue of duo_gemm) with out_kind
 * DUO_ACT_SPLIT, S <= 64, q_rows == S: the global patch attention in split-bf16 precision on tcgen05 / TMEM
 * (three bf16 UMMAs per product: hi*hi + hi*lo + lo*hi; fp32-grade result).
 * algo: 0 = auto, 1 = warp-per-(group,head) register/shuffle FMA kernel (any S <= 160),
 *       2 = warp-level tensor-core (mma.sync) kernel (bf16 in, bf16 out, 16 < S <= 96),
 *       3 = tcgen05 / TMEM kernel (bf16 in, bf16 out, 64 < S <= 96, q_rows == S: the 4-scale
 *           group size S = 86; auto picks it whenever it applies),
 *       4 = warp-per-(group, head) register kernel for S <= 8 (bf16 in, bf16 out: the 2-scale group size S = 6;
 *           mma.sync fragments in registers, softmax by quad shuffles; auto picks it whenever it applies).
 * q_rows: only the first q_rows query rows of every group are computed and `out` is the dense
 *       [num_groups * q_rows, D] matrix of those rows (q_rows = S: everything; q_rows = 1: the
 *       scale-token / CLS query only — all that the reference consumes after the LAST scale
 *       block, scale_attention.py:183-185, and after the last patch block, :341).
 */
int duo_group_attention(const void* qkv, int32_t in_kind, void* out, int32_t out_kind,
                        int64_t num_groups, int32_t S, int32_t num_heads, float scale,
                        int32_t algo, int32_t q_rows, duo_stream_t stream);

/*
 * Scale token row (s = 0) of the token tensor:  X[b,p,0,:] = tok[b,p,:] + pos0[:]
 * (tok strides in elements; 0,0 broadcasts the learned channel_token).
 * model_wo_extra_params.py:296-299 + scale_attention.py:331 ; model.py:322.
 */
int duo_fill_scale_token(float* X, const float* tok, int64_t tok_stride_b, int64_t tok_stride_p,
                         const float* pos0, int32_t B, int32_t P, int32_t S, int32_t D,
                         duo_stream_t stream);

/*
 * out[r,:] = in[r,:] + pos[r % S,:]  — `x + pos_embed_for_scale` for callers that hand a
 * ready-made token tensor to MultiscaleFormer / MultiscaleTransformer
 * (scale_attention.py:331 ; multi_vision_transformer.py:142-144).  in == out allowed.
 */
int duo_add_pos(const float* in, const float* pos, float* out, int64_t rows, int32_t S, int32_t D,
                duo_stream_t stream);

/*
 * Patch-stage input:  Z[b,0,:] = cls + pos[0];  Z[b,1+p,:] = X[b,p,0,:] + pos[1+p]
 * scale_attention.py:183-193 ; multiscale_attn.py:190-203.  out_kind BF16 or SPLIT.
 */
int duo_assemble_patch_tokens(const float* X, const float* cls, const float* pos, void* Z,
                              int32_t out_kind, int32_t B, int32_t P, int32_t S, int32_t D,
                              duo_stream_t stream);

/*
 * Classification head on one row per image: logits[b,:] = W * f(z_b) + bias, where
 * z_b = in[b*row_stride : +D] and f = LayerNorm(ln_gamma, ln_beta, eps) if ln_gamma != NULL.
 * scale_attention.py:341-344 (no norm) ; multi_vision_transformer.py:161-171 (norm, head).
 */
int duo_head(const float* in, int64_t row_stride, const float* ln_gamma, const float* ln_beta,
             float eps, const float* W, const float* bias, float* logits, int32_t B, int32_t D,
             int32_t num_classes, duo_stream_t stream);

/*
 * Channel-token branch (projection_head.py:152-268) on the GEMM kernel:
 * duo_im2col3x3: NHWC in [B,H,W,C] (bf16 / f16 / f32) -> bf16 [B*Ho*Wo, 9*C], column order (ky,kx,c),
 *   kernel 3, padding 1, stride 1 or 2 (Conv2d(k=3, s, p=1) == this + duo_gemm with the weight
 *   permuted to [N, ky, kx, c]).
 * duo_pool_to_slice: 2x2/2 max-pool (pool == 2) or copy (pool == 1) of an NHWC map into a channel
 *   slice of a wider NHWC tensor (`out` = first channel of the slice, rows ld_out apart): the
 *   MaxPool2d + torch.cat of model.py:279-286.
 */
int duo_im2col3x3(const void* in, int32_t in_kind, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                  int32_t stride, duo_stream_t stream);
int duo_pool_to_slice(const void* in, int32_t in_kind, void* out, int64_t ld_out, int32_t B, int32_t H,
                      int32_t W, int32_t C, int32_t pool, duo_stream_t stream);

/*
 * Trunk stem pooling: nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of torchvision's ResNet
 * (resnet50ssl.py:35-45 / torchvision resnet.py, between conv1 and layer1) on an NHWC map,
 * bf16 or f16 (`kind`), same type out: [B,H,W,C] -> [B,ceil(H/2),ceil(W/2),C].  C % 8 == 0.
 */
int duo_maxpool3x3s2(const void* in, int32_t kind, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                     duo_stream_t stream);

/*
 * Convolution of an NHWC 16-bit tensor as an implicit GEMM on tcgen05 / TMEM (no im2col matrix: the K loop walks the
 * filter taps, each operand tile is one 4-D TMA box load that applies the stride and the zero padding):
 *   out[b,ho,wo,n] = act( sum_{ky,kx,c} in[b, ho*s + ky - pad, wo*s + kx - pad, c] * weight[n, (ky,kx,c)] + bias[n]
 *                         (+ residual[b,ho,wo,n]) ),   pad = ksize / 2,  act = ReLU if relu else identity.
 * Replaces the cuDNN convolutions of the ResNet-50 trunk with BatchNorm folded into weight / bias
 * (torchvision Bottleneck: conv1 1x1, conv2 3x3 stride s, conv3 1x1 + identity / downsample 1x1 stride s; the producer
 * of the stage maps, model_wo_extra_params.py:214-224, model.py:213-223, resnet50ssl.py:35-45) and the 3x3 convolutions
 * of the channel-token branch (projection_head.py:152-268).
 * Cin % 64 == 0, Cout % 64 == 0, ksize 1 or 3, stride 1 or 2; in / weight fp16 (fp16 = 1) or bf16, out / residual fp16
 * (out_fp16 = 1) or bf16 — fp16 trunk maps can feed a convolution whose un-normalised output needs bf16's range; all
 * tensors 16-byte aligned.
 */
typedef struct duo_conv2d_args {
  const void* in;       /* NHWC [B, H, W, Cin]                                              */
  const void* weight;   /* [Cout, ksize*ksize*Cin], column order (ky, kx, c)                */
  const float* bias;    /* [Cout] or NULL                                                   */
  const void* residual; /* NHWC [B, Ho, Wo, Cout] added before the activation, or NULL      */
  void* out;            /* NHWC [B, Ho, Wo, Cout], Ho = (H + 2*pad - ksize) / stride + 1    */
  int32_t B, H, W, Cin, Cout;
  int32_t ksize, stride, relu, fp16, out_fp16;
  /* optional fused projection shortcut (torchvision Bottleneck.downsample: 1x1 convolution of the BLOCK INPUT with
   * stride stride2, summed into conv3's output): in2 NHWC [B, H2, W2, Cin2], same type as in; its weights are the LAST
   * Cin2 columns of `weight` ([Cout, ksize*ksize*Cin + Cin2]); (H2 - 1) / stride2 + 1 must equal Ho.  NULL: none. */
  const void* in2;
  int32_t H2, W2, Cin2, stride2;
} duo_conv2d_args;
int duo_conv2d(const duo_conv2d_args* args, duo_stream_t stream);

/*
 * Trunk stem (torchvision ResNet conv1: 7x7, stride 2, padding 3, 3 input channels, + folded bn1 + ReLU;
 * resnet50ssl.py:35-45 / model_wo_extra_params.py:214-224 child '0'..'2') on the same implicit-GEMM kernel.
 * duo_stem_pack: fp32 image x[b,c,y,x] (element strides given: NCHW or channels-last) times `scale` ->
 *   zero-padded row-pair tensor out [B, H + 8, W + 8, 8], fp16 or bf16: out[b,R,X,c] = pixel (R - 3, X - 3),
 *   out[b,R,X,4 + c] = pixel (R - 2, X - 3) for c < 3, zero outside the image and for c = 3.
 * duo_stem_conv7x7: packed -> out NHWC [B, H/2, W/2, Cout]; weight [Cout, 4 * 64] with the coefficient of filter tap
 *   (ky, kx, c) in column (ky / 2) * 64 + kx * 8 + (ky % 2) * 4 + c (ky, kx < 7, c < 3; every other column zero): one
 *   64-wide K block = 8 pixels x two filter rows x 4 channels = 128 contiguous bytes of the packed tensor, loaded by ONE
 *   TMA box of overlapping windows.  H, W even, Cout % 64 == 0.
 */
int duo_stem_pack(const float* x, int64_t stride_b, int64_t stride_c, int64_t stride_h, int64_t stride_w, float scale,
                  void* out, int32_t fp16, int32_t B, int32_t H, int32_t W, duo_stream_t stream);
int duo_stem_conv7x7(const void* packed, const void* weight, const float* bias, void* out, int32_t B, int32_t H,
                     int32_t W, int32_t Cout, int32_t relu, int32_t fp16, duo_stream_t stream);

/* fp32 [rows, cols] (leading dim ld) -> bf16 / split bf16 (contiguous). */
int duo_convert(const float* in, int64_t ld, void* out, int32_t out_kind, int64_t rows,
                int32_t cols, duo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DUOFORMER_SM100_H_ */
