// Training step of minGPT (BASELINE config 4): teacher-forced forward with dropout,
// backward through every layer, fused AdamW.  reference: Lit_minGPT.training_step / shared_step
// transformer/minGPT.py:413-422, configure_optimizers :618-665, DDP wrapping GPT_VAE_train.py:172-174.
//
// All dense contractions (forward, dgrad, wgrad) run on the tcgen05 GEMM of gemm_tc.cu.  That kernel contracts two
// K-major operands, D[M,N] = A[M,K] B[N,K]^T, so
//   forward  Y  = X W^T          : A = X [R, in],   B = W   [out, in]
//   dgrad    dX = dY W           : A = dY [R, out], B = W^T [in, out]      (transposed bf16 weight copies, refreshed by the optimizer)
//   wgrad    dW = dY^T X         : A = dY^T [out, Rpad], B = X^T [in, Rpad] (activations / gradients transposed on the fly)
// The non-GEMM kernels live in gpt_train_kernels.cu.
#pragma once
#include "gpt_kernels.cuh"

namespace mgv {

// dropout: keep(element) = hash(seed, stream, index) >= p (24-bit threshold); thresh24 == 0 <=> no dropout
struct DropCfg {
  uint32_t seed_lo, seed_hi, thresh24;
};
constexpr uint32_t DROP_STREAM_EMBD = 1;
inline uint32_t drop_stream_attn(int layer) { return 16u + 4u * static_cast<uint32_t>(layer); }
inline uint32_t drop_stream_resid_attn(int layer) { return 17u + 4u * static_cast<uint32_t>(layer); }
inline uint32_t drop_stream_resid_mlp(int layer) { return 18u + 4u * static_cast<uint32_t>(layer); }
DropCfg make_drop(float p, unsigned long long seed);
float drop_inv_keep(const DropCfg& d);

// one tensor of the flat parameter / gradient buffers
struct AdamSeg {
  long long offset, numel;
  int decay;                  // 1: weight decay applies (Linear weights), 0: biases, LayerNorm, embeddings (minGPT.py:629-647)
  float* dst_f32;             // refreshed fp32 copy inside the handle (or null)
  __nv_bfloat16* dst_bf16;    // refreshed bf16 copy inside the handle (or null)
};
constexpr int ADAM_CHUNK = 16384;   // elements per CTA of the fused AdamW kernel

enum { GRAD_PREP_DROP = 0, GRAD_PREP_GELU = 1, GRAD_PREP_COPY = 2 };

// keep flags (1 / 0) of elements 0..n-1 of a dropout stream (tests: rebuild the masks the kernels regenerate)
int train_drop_mask(const DropCfg& dc, unsigned stream_id, long long n, unsigned char* out, cudaStream_t s);
int train_embed(const long long* idx, int B, int T, int t, const long long* cls, const float* embedder, int m,
                const float* tok_emb, const float* pos_emb, int C, int vocab, int class_size, float* x_out, int* err_flag,
                const DropCfg& dc, cudaStream_t s);
int train_resid_dropout(const float* resid, const float* branch, long long n, float* x_out, const DropCfg& dc,
                        unsigned stream_id, cudaStream_t s);
int train_gelu_fwd(const __nv_bfloat16* hpre, long long n, __nv_bfloat16* h, cudaStream_t s);
// src [R,N] (fp32 for DROP, bf16 otherwise) -> g [R,N] bf16 (not for COPY), gT [N,Rpad] bf16, db [N] += column sums (optional)
int train_grad_prep(int mode, const void* src, const __nv_bfloat16* aux, int R, int N, int Rpad, __nv_bfloat16* g,
                    __nv_bfloat16* gT, float* db, const DropCfg& dc, unsigned stream_id, cudaStream_t s);
int train_transpose(const __nv_bfloat16* src, int R, int N, long long ld_src, int Rpad, __nv_bfloat16* dst, cudaStream_t s);
int train_layernorm_bwd(const float* dy, const float* x, const float* gamma, int rows, int C, float* dx_io, bool accumulate,
                        float* dgamma, float* dbeta, cudaStream_t s);
int train_ce(const float* logits, const long long* targets, long long rows, int V, float* loss_sum,
             __nv_bfloat16* dlogits, int* err_flag, cudaStream_t s);
int train_embed_bwd(const float* dx, const long long* idx, int B, int T, int t, const long long* cls, int m, int C,
                    float* d_tok, float* d_pos, float* d_embedder, const DropCfg& dc, cudaStream_t s);
// y [B*T, C] bf16, lse2 [B*nh, T] fp32 (log2 domain, scale folded)
int train_attn_fwd(const __nv_bfloat16* qkv, int B, int T, int nh, int n_unmasked, __nv_bfloat16* y, float* lse2,
                   const DropCfg& dc, unsigned stream_id, cudaStream_t s);
// dqkv [B*T, 3C] bf16 (all three slots written); delta [B*nh, T] scratch
int train_attn_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* dy, const float* lse2, int B, int T,
                   int nh, int n_unmasked, __nv_bfloat16* dqkv, float* delta, const DropCfg& dc, unsigned stream_id,
                   cudaStream_t s);
int train_adamw(float* p, const float* g, float* m, float* v, const AdamSeg* d_segs, const int2* d_chunks, int n_chunks,
                float lr, float beta1, float beta2, float eps, float wd, long long step, float grad_scale, cudaStream_t s);

// ---- handle-level API (gpt_train.cu)
struct Gpt;
long long gpt_train_numel(Gpt* g);
int gpt_train_layout(Gpt* g, const char* name, long long* offset, long long* numel, int* decay);
int gpt_train_bind(Gpt* g, float* flat_params, float* flat_grads, cudaStream_t s);
int gpt_train_forward(Gpt* g, const long long* idx, int B, int t, const long long* cls, int m, const long long* targets,
                      float p_embd, float p_resid, float p_attn, unsigned long long seed, float* loss_out, cudaStream_t s);
int gpt_train_backward(Gpt* g, int layer_hi, int layer_lo, cudaStream_t s);
int gpt_train_adamw(Gpt* g, float* m, float* v, float lr, float beta1, float beta2, float eps, float weight_decay,
                    long long step, float grad_scale, cudaStream_t s);

}  // namespace mgv
