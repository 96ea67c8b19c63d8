// tcgen05 / TMEM / TMA GEMM used by every dense contraction on the token path:
//   * minGPT Linear layers (prefill and the per-token decode step, where it is a
//     weight-streaming kernel: M = batch, split-K over the grid),
//   * the VQVAE 1x1 convolutions (plain GEMM over NHWC pixels),
//   * the VQVAE 3x3 convolutions as implicit GEMM: the A operand is fetched by TMA from
//     the NHWC activation tensor with one shifted 4-D box per filter tap (zero fill at
//     the image border = the convolution's zero padding; element stride 2 = stride-2 conv).
// D[M,N] = A[M,K] * B[N,K]^T, A/B bf16 K-major, fp32 accumulation in TMEM.
#pragma once
#include <stdlib.h>
#include "mgv_sm100.cuh"

namespace mgv {

enum GemmEpilogue : int {
  EPI_BF16 = 0,        // out_bf16 = acc + bias
  EPI_BF16_GELU = 1,   // out_bf16 = gelu_erf(acc + bias)
  EPI_F32 = 2,         // out_f32 = acc + bias
  EPI_F32_RESID = 3,   // out_f32 = resid_f32 + acc + bias        (resid may alias out)
  EPI_F32_ATOMIC = 4,  // out_f32 += acc (+ bias from split 0)     (split-K; out pre-initialised)
  EPI_BF16_RESID = 5,  // out_bf16 = resid_bf16 + acc + bias       (conv residual blocks)
};

enum GemmAMode : int {
  A_PLAIN = 0,    // A is a row-major [M,K] matrix
  A_CONV3x3 = 1,  // A rows are output pixels of a 3x3 convolution over NHWC input
};

struct GemmArgs {
  // operands
  const void* A = nullptr;  // bf16: [M,K] (lda elements) or NHWC input [N,Hin,Win,Cin]
  const void* B = nullptr;  // bf16 weights [N,K] row-major (K = 9*Cin for conv, tap-major)
  int M = 0, N = 0, K = 0;
  int64_t lda = 0;          // elements; 0 -> K
  // epilogue
  int epi = EPI_BF16;
  const float* bias = nullptr;  // [N] or null
  void* out = nullptr;
  const void* resid = nullptr;
  int64_t ldo = 0;  // elements; 0 -> N
  // tiling
  int bn = 128;     // 32 / 64 / 128 / 256
  int split_k = 1;  // EPI_F32_ATOMIC only when > 1
  int max_stages = 0;
  // conv geometry (A_CONV3x3): output HxW, input Hin x Win, stride 1 or 2
  int a_mode = A_PLAIN;
  int n_img = 0, H = 0, W = 0, Hin = 0, Win = 0, Cin = 0, stride = 1, pad = 1;
  // generalisations used by the phase form of Upsample (nearest 2x + 3x3 conv == four 2x2 convs on the low-res tensor):
  // taps per filter row (K = taps_y * taps_x * Cin), a y padding different from `pad` (-1: same), and a strided output
  // placement: pixel (y, x) of the H x W grid goes to (y*out_scale + out_oy, x*out_scale + out_ox) of an
  // (H*out_scale) x (W*out_scale) output image.  GroupNorm partial sums of tile (img, ty, tx) go to slot
  // img * gn_tiles_img + gn_tile_base + ty * tiles_x + tx (0: tiles per image of this launch, base 0).
  int taps_x = 3, pad_y = -1, out_scale = 1, out_oy = 0, out_ox = 0, gn_tiles_img = 0, gn_tile_base = 0;
  // GroupNorm statistics fused into the epilogue (conv): per (image, channel-group of
  // `gn_group_ch` output channels) sum and sum of squares of the bf16-rounded outputs.
  // Every CTA stores (no atomics -> deterministic) its partial sums to
  // gn_sum[(img * tiles_per_image + tile) * (N / gn_group_ch) * 2 + group * 2 + {0: sum, 1: sumsq}];
  // vqvae_gn_finalize folds the tiles in a fixed order.
  float* gn_sum = nullptr;
  int gn_group_ch = 0;
  // swap-AB (decode): A = weights [M = out features, K], B = activations [N = batch rows, K]; the
  // accumulator tile is written TRANSPOSED: out[n * ldo + m] (ldo defaults to M), bias is per m.
  // N need not be a multiple of 32 (TMA zero-fills the missing activation rows).
  bool transpose_out = false;
  // launch
  bool pdl = false;
  bool weights_evict_first = false;  // L2 evict-first hint on the weight operand (B, or A when transpose_out)
  cudaStream_t stream = nullptr;
};

int gemm_bf16_tc(const GemmArgs& a);

// Width (in output pixels) of a conv M tile: the 128 rows of a tile are a wb x (128 / wb) pixel rectangle, and the choice
// among 128 / 64 / 32 / 16 that covers the H x W image with the fewest padded pixels wins (ties: the wider one).  E.g.
// 80 x 848 is covered exactly by 16 x 8 rectangles while 128 x 1 strips pad every row to 896 (5.7 % wasted MMA work),
// and 40 x 424 pads to 512 with strips (20.8 %) but to 432 with 16 x 8 rectangles (1.9 %).
inline int conv_tile_width(int H, int W) {
  static const bool legacy = getenv("MGV_CONV_WB_LEGACY") != nullptr;   // first version: the widest strip that fits
  if (legacy) {
    int wb = 16;
    while (wb < 128 && wb < W) wb *= 2;
    return wb;
  }
  int best = 128;
  long long best_area = -1;
  for (int wb = 128; wb >= 16; wb /= 2) {
    const int hb = 128 / wb;
    const long long area = static_cast<long long>((W + wb - 1) / wb) * wb * ((H + hb - 1) / hb) * hb;
    if (best_area < 0 || area < best_area) {
      best_area = area;
      best = wb;
    }
  }
  return best;
}

// Decode-step GEMM without split-K (gemm_decode_fullk.cu): a CTA = 128 weight rows x 32 sequences x all of K;
// epi in {EPI_BF16_GELU, EPI_F32, EPI_F32_RESID}; W bf16 [Nw, K], X bf16 [B, K], K % 256 == 0.
int gemm_decode_fullk(const void* W, int Nw, int K, const void* X_bf16, int B, const float* bias, int epi, void* out,
                      const void* resid, long long ldo, bool pdl, cudaStream_t stream);

// ---- decode-step GEMMs that absorb the LayerNorm / GELU stage in front of them (gemm_decode_fold.cu)
// (LnFold: mgv_common.cuh)
enum FoldMode : int { FOLD_LN = 0, FOLD_GELU = 1 };
// mode FOLD_LN  : out[b, n] += sum_k W[n,k] bf16(gamma[k] src[b,k]);  stats_out[z][b] = partial row statistics of src
// mode FOLD_GELU: out[b, n] += sum_k W[n,k] bf16(gelu(rstd_b (src[b,k] - mu_b in.sw[k]) + in.bp[k])) (+ bias[n])
// W bf16 [Nw, K]; src fp32 [B, K]; out fp32 [B, ldo] (holds zeros / the residual); kbps = 64-wide k-blocks per CTA (<= 4);
// bn = sequences per CTA (32 or 64).
int gemm_decode_fold(int mode, const void* W, int Nw, int K, const float* src, int B, const float* gamma,
                     float2* stats_out, int stats_stride, const LnFold* in, const float* bias, float* out, long long ldo,
                     int kbps, int staging_warps, int bn, bool pdl, cudaStream_t stream);
// sw[n] = sum_k W[n,k] gamma[k];  bp[n] = sum_k W[n,k] beta[k] + bias[n] (bias may be null)
int gpt_fold_prepare(const void* W, int N, int K, const float* gamma, const float* beta, const float* bias, float* sw,
                     float* bp, cudaStream_t stream);

// out[b, n] = rstd_b * (out[b, n] - mu_b * f.sw[n]) + f.bp[n]: the consumer-side half of FOLD_LN as a stand-alone kernel (tests)
int gpt_fold_apply(float* out, int B, int N, const LnFold& f, cudaStream_t stream);

// SIMT fp32 reference of the same contract (tests / on-device cross-checks only).
int gemm_bf16_ref(const GemmArgs& a);

}  // namespace mgv
