// Non-GEMM kernels of the VQVAE encoder / decoder (NHWC bf16 activations).
// reference: vqvae/big_model_attn_gan.py (ResnetBlock :114-135, Normalize :139-140,
// nonlinearity :164-166, Upsample :182-186, AttnBlock :425-450, Decoder :361-392).
#pragma once
#include "mgv_common.cuh"

namespace mgv {

// conv weight (Cout, Cin, KH, KW) fp32 -> [Cout][KH*KW][Cin] bf16 (K-major rows for the GEMM)
int vqvae_repack_conv_weight(const float* src, int Cout, int Cin, int KH, int KW, __nv_bfloat16* dst, cudaStream_t s);
// phase weights of Upsample (nearest 2x + 3x3 conv) as four 2x2 convs on the low-res tensor:
// src fp32 OIHW (Cout, Cin, 3, 3) -> dst bf16 [py*2+px][Cout][ty*2+tx][Cin]
int vqvae_upsample_phase_weights(const float* src, int Cout, int Cin, __nv_bfloat16* dst, cudaStream_t s);

// table[k][:] = codebook[k] @ Wpq^T + bpq   (get_codebook_entry :56-71 fused with post_quant_conv :611)
int vqvae_build_gather_table(const float* codebook, const float* wpq /*(Cout,Cin) fp32*/, const float* bpq, int K,
                             int Cin, int Cout, __nv_bfloat16* table, cudaStream_t s);
// out[n, :] = table[idx[n], :]  (NHWC bf16 rows); bad_flag set if an index is outside [0,K)
int vqvae_gather_rows(const long long* idx, const __nv_bfloat16* table, long long n, int C, int K, __nv_bfloat16* out,
                      int* bad_flag, cudaStream_t s);

// layout changes at the path boundary
int vqvae_nchw_f32_to_nhwc_bf16(const float* in, int N, int C, int HW, __nv_bfloat16* out, cudaStream_t s);
int vqvae_nhwc_f32_to_nchw_f32(const float* in, int N, int C, int HW, float* out, cudaStream_t s);

// GroupNorm(32 groups, eps 1e-6) statistics as per-tile partial sums: part[(n * tiles + tile) * 64 + g * 2 + {sum, sumsq}].
// Producers: the conv epilogue (gemm_tc.cu, one slot per CTA) or vqvae_gn_stats (one slot per chunk; returns the
// number of chunks per image in *tiles_out).  No atomics: the result is deterministic.
int vqvae_gn_stats(const __nv_bfloat16* x, int N, int HW, int C, float* part, int* tiles_out, cudaStream_t s);
// folds the tiles in a fixed order: mr[n][g] = {mean, rstd}; count = HW * C/32
int vqvae_gn_finalize(const float* part, int N, int tiles, float count, float* mr /*[N,32,2]*/, cudaStream_t s);
// y = gn(x) * gamma + beta, optionally followed by swish (x * sigmoid(x)); mr from vqvae_gn_finalize
int vqvae_gn_apply(const __nv_bfloat16* x, const float* mr, const float* gamma, const float* beta, int N, int HW, int C,
                   int do_swish, __nv_bfloat16* y, cudaStream_t s);

// nearest-neighbour 2x upsample (F.interpolate scale_factor=2 mode="nearest", :183)
int vqvae_upsample2x(const __nv_bfloat16* x, int N, int H, int W, int C, __nv_bfloat16* y, cudaStream_t s);

// AttnBlock core (:434-446): qkv bf16 [N*T, 3C] = [q|k|v] -> o bf16 [N*T, C]; single head, scale C^-0.5
int vqvae_spatial_attention(const __nv_bfloat16* qkv, int N, int T, int C, __nv_bfloat16* o, cudaStream_t s);

// Decoder tail (:389-391): out[n,0,y,x] = conv3x3(swish(gn(h)), w (1,C,3,3)) + b ; h NHWC bf16, out fp32
int vqvae_norm_swish_conv_out(const __nv_bfloat16* h, const float* mr, const float* gamma, const float* beta,
                              const float* w /*[9][C] fp32*/, const float* bias, int N, int H, int W, int C, float* out,
                              cudaStream_t s);

// Encoder head (:261): out NHWC bf16 [N,H,W,Cout] = conv3x3(mel (N,1,H,W) fp32, w [Cout][9] fp32) + b
int vqvae_conv_in_1ch(const float* mel, const float* w, const float* bias, int N, int H, int W, int Cout,
                      __nv_bfloat16* out, cudaStream_t s);

}  // namespace mgv
