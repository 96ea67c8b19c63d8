// Non-GEMM kernels of the minGPT path (embedding, LayerNorm, attention, GELU, fused
// final-LN + head + top-k + sampling).  reference: transformer/minGPT.py:45-199, 287-360.
#pragma once
#include "mgv_common.cuh"

namespace mgv {

constexpr int GPT_HEAD_DIM = 64;   // n_embd / n_head for every reference config
constexpr int GPT_MAX_T = 288;     // >= block_size (266), multiple of 32

// x[b,t,:] = (t < m ? prefix : tok_emb[idx[b,t-m]]) + pos_emb[t]            (minGPT.py:170-180)
// prefix is either prefix_emb[b,t,:] (float embeddings) or embedder[cls[b]] (GPTClass, :210)
// Rows p = p_off .. p_off+R-1 of every sequence are produced: x_out is [B, R, C].
int gpt_embed(const long long* idx, int B, int R, int p_off, int idx_ld, const float* prefix_emb, const long long* cls,
              const float* embedder, int m, const float* tok_emb, const float* pos_emb, int C, int vocab,
              int class_size, float* x_out, int* err_flag, cudaStream_t s, bool pdl);

// LayerNorm (eps 1e-5, affine) fp32 rows -> bf16 rows (zero_buf / zero_count: unused, kept for ABI stability).
int gpt_layernorm(const float* x, const float* w, const float* b, int rows, int C, __nv_bfloat16* out, float* zero_buf,
                  long long zero_count, cudaStream_t s, bool pdl);

// Causal (or prefix-unmasked) self-attention over a whole sequence (prefill / teacher forcing).
// qkv: bf16 [B*T, 3C] = [q | k | v]; y: bf16 [B*T, C];
// att (optional): fp32 [B, nh, att_T, att_T], PRE-ZEROED by the caller; only unmasked entries are written;
// kcache/vcache (optional): bf16 [B, nh, Tmax, 64] receive K and V of positions 0..T-1.
int gpt_attention_prefill(const __nv_bfloat16* qkv, int B, int T, int nh, int n_unmasked, __nv_bfloat16* y, float* att,
                          int att_T, __nv_bfloat16* kcache, __nv_bfloat16* vcache, int Tmax, cudaStream_t s);
// the mma.sync kernel (every mask / output option)
int gpt_attention_prefill_mma(const __nv_bfloat16* qkv, int B, int T, int nh, int n_unmasked, __nv_bfloat16* y, float* att,
                              int att_T, __nv_bfloat16* kcache, __nv_bfloat16* vcache, int Tmax, cudaStream_t s);
// The same contract on tcgen05 / TMEM / TMA (attn_prefill_tc.cu) for the plain causal mask without the attention-map
// output and without the KV-cache fill; gpt_attention_prefill routes to it when `supported(T)`.
bool gpt_attention_prefill_tc_supported(int T);
// trace (optional, diagnostics): 48 clock64 stamps of the middle CTA (tools/attn_trace.py)
int gpt_attention_prefill_tc(const __nv_bfloat16* qkv, int B, int T, int nh, __nv_bfloat16* y, cudaStream_t s,
                             long long* trace = nullptr);

// One decode position: q,k,v fp32 [B, 3C] for position *pos_ptr; appends k,v to the cache and
// attends over positions 0..pos.  att_rows (optional): fp32 [B, nh, Tatt, Tatt] pre-zeroed; entries 0..pos of row pos are written.
// zero_consumed: the q,k,v accumulators are cleared after they are read (ready for the next split-K reduction).
// zero_buf / zero_count: optional second fp32 buffer cleared cooperatively (the FC1 split-K accumulator).
// fold: q, k, v are raw accumulators of a FOLD_LN GEMM (gemm_decode_fold.cu); the LayerNorm is applied here.
int gpt_attention_decode(float* qkv32, int B, int nh, const int* pos_ptr, __nv_bfloat16* kcache,
                         __nv_bfloat16* vcache, int Tmax, __nv_bfloat16* y, float* att_rows, int Tatt, bool zero_consumed,
                         float* zero_buf, long long zero_count, cudaStream_t s, bool pdl, const LnFold* fold = nullptr);

// h = gelu_erf(h32) as bf16 (split-K FC1 path)
int gpt_gelu_bf16(float* h32, long long n, __nv_bfloat16* out, bool zero_consumed, cudaStream_t s, bool pdl);

struct SampleArgs {
  float* logits_acc;       // [B, V] fp32: head logits of this position (split-K accumulator; cleared here after use)
  int B, C, V;
  int row0;                // index of this group's first sequence in the whole batch (Philox counter = row0 + b)
  float temperature;
  int top_k;               // 0 = no top-k
  int do_sample;           // 0 = greedy (torch.topk(probs, 1)), 1 = multinomial
  const unsigned long long* seed_ptr;   // device: Philox key (read per step, so a captured graph can be reused)
  int* pos_ptr;            // device: position of the row just processed; incremented here
  long long* tokens;       // [B, tokens_ld]; the sampled token goes to slot pos + 1 - m
  int tokens_ld;
  int m;                   // prefix length (cond_size)
  const float* tok_emb;
  const float* pos_emb;
  int block_size;
  float* x_next;           // [B, C]: embedding of the sampled token at position pos+1
  float* logits_out;       // optional [steps, logits_step_stride]: logits of every step (after temperature, before top-k),
                           // step index = position - logits_pos0, row b at offset (row0 + b) * V
  long long logits_step_stride;
  int logits_pos0;
  unsigned int* done_counter;  // device scratch (zero-initialised) used to advance *pos_ptr once per step
  LnFold fold;             // folded ln_f (stats == nullptr: logits_acc already holds the logits)
};
int gpt_sample_step(const SampleArgs& a, cudaStream_t s, bool pdl);

// loss[r] = logsumexp(logits[r, :V]) - logits[r, targets[r]]; a target outside [0, V) sets *err_flag = 2
int gpt_ce_rows(const float* logits, const long long* targets, long long rows, int V, long long ld, float* loss,
                int* err_flag, cudaStream_t s);

}  // namespace mgv
