// Private definition of the minGPT handle, shared by gpt.cu (inference: prefill, decode loop) and gpt_train.cu
// (teacher-forced training step).  Not part of the C ABI.
#pragma once
#include <vector>
#include "gpt.cuh"
#include "gpt_kernels.cuh"

namespace mgv {

struct GptTrain;   // training-step state (gpt_train.cu); null until mgv_gpt_train_bind

struct GptLayer {
  float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  __nv_bfloat16 *wqkv, *wproj, *wfc1, *wfc2;
  float *bqkv, *bproj, *bfc1, *bfc2;
  // folded-LayerNorm vectors of the decode chain (gemm_decode_fold.cu): sw = W gamma, bp = W beta + b
  float *sw_qkv, *bp_qkv, *sw_fc1, *bp_fc1;
};

// split-K factors of the decode-step GEMMs (swap-AB: 128 weight rows per CTA x split-K slices)
struct DecodeTiles {
  int qkv_split = 4;    // 24 row tiles x 4 (x 2 sequence halves at batch 64) = 192 CTAs
  int proj_split = 16;  //  8 row tiles x 16 = 128 CTAs
  int fc1_split = 4;    // 32 row tiles x 4  = 128 CTAs
  int fc2_split = 16;   //  8 row tiles x 16 = 128 CTAs
  int head_split = 16;  //  1 row tile  x 16 (vocab 128)
};

constexpr int FOLD_MAX_PARTS = 64;

struct Gpt {
  GptConfig cfg;
  int C, L, nh, V, Vout, Tmax;
  // parameters
  void* slab = nullptr;
  size_t slab_bytes = 0;
  float *tok_emb, *pos_emb, *embedder, *lnf_w, *lnf_b;
  __nv_bfloat16* whead;
  std::vector<GptLayer> layers;
  std::vector<unsigned char> loaded;  // per tensor
  int n_tensors = 0;
  // workspaces
  int ws_rows = 0;  // prefill rows capacity
  float* x = nullptr;
  __nv_bfloat16 *ln = nullptr, *qkv = nullptr, *y = nullptr, *h = nullptr;
  int dec_B = 0;  // decode batch capacity
  float *dx = nullptr, *dqkv32 = nullptr, *dh32 = nullptr, *dlogits = nullptr;
  __nv_bfloat16 *dln = nullptr, *dy = nullptr, *dh = nullptr, *kv = nullptr;
  // [2]=err flag, [4..5]=Philox seed (u64), [8+2c]=position of sequence group c, [9+2c]=its done counter
  int* d_state = nullptr;
  long long* dtokens = nullptr;  // [dec_B, block_size] token buffer the sampler writes (stable address for the graph)
  // cached decode-step graph (valid while the key matches and the workspaces are not reallocated)
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  struct { int B, m, top_k, do_sample; float temperature; long long per_step; } graph_key = {0, 0, 0, 0, 0.f, 0};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // The decode step of one position is a chain of dependent ~4 us kernels that leaves most of the GPU idle, so the
  // B sequences are split into `groups` independent groups whose chains run concurrently (parallel branches of the
  // step graph) and fill each other's bubbles.  Rows of different sequences never interact, so results do not change.
  static constexpr int MAX_GROUPS = 8;
  int groups = 1;   // MGV_DECODE_GROUPS: measured 935 (2 groups) vs 1016 us per position, but every chain still pays the
                    // full per-stage latency, and the default stays one group (deterministic launch order, simpler graph)
  // Folded decode chain (default): LayerNorm / GELU are applied by the consumer of each split-K accumulator, so a
  // block is 5 dependent kernels (QKV, attention, proj, FC1, FC2) instead of 7.  MGV_DECODE_FOLD=0 selects the
  // separate-LayerNorm chain (also used for shapes the fold kernels do not cover).
  bool use_fold = true;
  bool fold_dirty = true;          // sw / bp vectors must be recomputed (a weight was loaded)
  float *sw_head = nullptr, *bp_head = nullptr;
  float2 *stats1 = nullptr, *stats2 = nullptr;   // [FOLD_MAX_PARTS][dec_B] partial LayerNorm statistics (ln1 / ln_f, ln2)
  int fold_sw = 4, fold_sw_gelu = 8;             // staging warps of the fold GEMMs (MGV_FOLD_SW=a,b)
  int fold_bn = 32, fold_bn2 = 32;               // sequences per CTA of the FOLD_LN / FOLD_GELU GEMMs (MGV_FOLD_BN=a,b)
  cudaStream_t gstream[MAX_GROUPS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_GROUPS] = {};
  bool pdl = false;
  DecodeTiles tiles;
  long long launches = 0;  // kernels launched by the last forward / generate call
  bool deterministic = false, saved_fold = true;   // gpt_set_deterministic
  DecodeTiles saved_tiles;
  int saved_groups = 1;
  float* step_logits = nullptr;   // one-shot request (gpt_set_step_logits): per-step logits of the next generate call
  GptTrain* train = nullptr;      // owned; released by gpt_destroy through gpt_train_release

  __nv_bfloat16* kcache(int l) const {
    return kv + (static_cast<size_t>(l) * 2) * dec_B * nh * Tmax * GPT_HEAD_DIM;
  }
  __nv_bfloat16* vcache(int l) const {
    return kv + (static_cast<size_t>(l) * 2 + 1) * dec_B * nh * Tmax * GPT_HEAD_DIM;
  }
};


void gpt_train_release(Gpt* g);   // gpt_train.cu

}  // namespace mgv
