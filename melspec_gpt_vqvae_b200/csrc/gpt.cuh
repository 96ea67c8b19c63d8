// minGPT handle (see gpt.cu).
#pragma once
#include "mgv_common.cuh"

namespace mgv {

struct GptConfig {
  int vocab_size;
  int block_size;
  int n_layer;
  int n_head;
  int n_embd;
  int class_size;   // 0: no class embedder (plain GPT)
  int n_unmasked;   // prefix block left unmasked (minGPT.py:67-68)
  int head_out;     // 0: vocab_size; else `last_linear` (minGPT.py:144-149)
};

struct Gpt;

int gpt_create(const GptConfig* cfg, Gpt** out);
int gpt_destroy(Gpt* g);
int gpt_load_weight(Gpt* g, const char* name, const float* src, long long numel, cudaStream_t s);
int gpt_forward(Gpt* g, const long long* idx, int B, int t, const float* prefix_emb, const long long* cls, int m,
                float* logits_out, float* att_out, cudaStream_t s);
int gpt_generate(Gpt* g, const long long* x0, int B, int t0, const float* prefix_emb, const long long* cls, int m,
                 int steps, float temperature, int do_sample, int top_k, unsigned long long seed, long long* x_out,
                 float* att_out, int use_graph, cudaStream_t caller);
int gpt_cross_entropy(Gpt* g, const float* logits, const long long* targets, long long rows, int V, float* loss,
                      cudaStream_t s);
long long gpt_last_launches(const Gpt* g);
// one-shot: the NEXT gpt_generate call also writes the logits of every step (after temperature, before top-k) to
// buf [steps, B, V] fp32 (device, caller-owned); pass nullptr to cancel
int gpt_set_step_logits(Gpt* g, float* buf);
// on != 0: decode without split-K reductions (bit-reproducible run to run, slower); 0 restores the default schedule
int gpt_set_deterministic(Gpt* g, int on);

}  // namespace mgv
