// minGPT handle: packed bf16 weights, KV cache, prefill (teacher-forced forward) and the
// CUDA-graph decode loop.  reference: transformer/minGPT.py:121-212 (GPT, GPTClass),
// :293-360 (Lit_minGPT.sample).
#include <string>
#include <vector>
#include <stdlib.h>
#include "gemm_tc.cuh"
#include "gpt_kernels.cuh"
#include "gpt.cuh"
#include "gpt_impl.cuh"

namespace mgv {

namespace {

__global__ void cvt_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}

int cvt_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t s) {
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  if (blocks < 1) blocks = 1;
  cvt_f32_bf16_kernel<<<blocks, 256, 0, s>>>(in, out, n);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int pick_bn(int N) {
  if (N % 256 == 0 && !getenv("MGV_NO_BN256")) return 256;   // 128x256 tiles: fewer L2 bytes per FLOP
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 32;
}

}  // namespace

namespace {

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct Carver {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align256(n * sizeof(T));
    return p;
  }
};

void carve_params(Gpt* g, char* base, size_t* total) {
  Carver c{base};
  const size_t C = g->C;
  g->tok_emb = c.take<float>(static_cast<size_t>(g->V) * C);
  g->pos_emb = c.take<float>(static_cast<size_t>(g->cfg.block_size) * C);
  g->embedder = c.take<float>(static_cast<size_t>(g->cfg.class_size > 0 ? g->cfg.class_size : 1) * C);
  g->lnf_w = c.take<float>(C);
  g->lnf_b = c.take<float>(C);
  g->whead = c.take<__nv_bfloat16>(static_cast<size_t>(g->Vout) * C);
  g->sw_head = c.take<float>(g->Vout);
  g->bp_head = c.take<float>(g->Vout);
  g->layers.resize(g->L);
  for (int l = 0; l < g->L; ++l) {
    GptLayer& y = g->layers[l];
    y.ln1_w = c.take<float>(C);
    y.ln1_b = c.take<float>(C);
    y.ln2_w = c.take<float>(C);
    y.ln2_b = c.take<float>(C);
    y.wqkv = c.take<__nv_bfloat16>(3 * C * C);
    y.bqkv = c.take<float>(3 * C);
    y.wproj = c.take<__nv_bfloat16>(C * C);
    y.bproj = c.take<float>(C);
    y.wfc1 = c.take<__nv_bfloat16>(4 * C * C);
    y.bfc1 = c.take<float>(4 * C);
    y.wfc2 = c.take<__nv_bfloat16>(4 * C * C);
    y.bfc2 = c.take<float>(C);
    y.sw_qkv = c.take<float>(3 * C);
    y.bp_qkv = c.take<float>(3 * C);
    y.sw_fc1 = c.take<float>(4 * C);
    y.bp_fc1 = c.take<float>(4 * C);
  }
  *total = c.off;
}

int ensure_prefill_ws(Gpt* g, int rows) {
  if (rows <= g->ws_rows) return MGV_OK;
  cudaFree(g->x); cudaFree(g->ln); cudaFree(g->qkv); cudaFree(g->y); cudaFree(g->h);
  g->x = nullptr; g->ln = g->qkv = g->y = g->h = nullptr;
  g->ws_rows = 0;
  const size_t C = g->C, R = rows;
  MGV_CHECK_CUDA(cudaMalloc(&g->x, R * C * 4));
  MGV_CHECK_CUDA(cudaMalloc(&g->ln, R * C * 2));
  MGV_CHECK_CUDA(cudaMalloc(&g->qkv, R * 3 * C * 2));
  MGV_CHECK_CUDA(cudaMalloc(&g->y, R * C * 2));
  MGV_CHECK_CUDA(cudaMalloc(&g->h, R * 4 * C * 2));
  g->ws_rows = rows;
  return MGV_OK;
}

int ensure_decode_ws(Gpt* g, int B) {
  if (B <= g->dec_B) return MGV_OK;
  cudaFree(g->dx); cudaFree(g->dqkv32); cudaFree(g->dh32); cudaFree(g->dln); cudaFree(g->dy); cudaFree(g->dh);
  cudaFree(g->kv); cudaFree(g->dlogits); cudaFree(g->dtokens); cudaFree(g->stats1); cudaFree(g->stats2);
  g->dtokens = nullptr;
  g->stats1 = g->stats2 = nullptr;
  if (g->graph_exec) { cudaGraphExecDestroy(g->graph_exec); g->graph_exec = nullptr; }
  if (g->graph) { cudaGraphDestroy(g->graph); g->graph = nullptr; }
  g->dx = g->dqkv32 = g->dh32 = g->dlogits = nullptr;
  g->dln = g->dy = g->dh = g->kv = nullptr;
  g->dec_B = 0;
  const size_t C = g->C, R = B;
  MGV_CHECK_CUDA(cudaMalloc(&g->dx, R * C * 4));
  MGV_CHECK_CUDA(cudaMalloc(&g->dqkv32, R * 3 * C * 4));
  MGV_CHECK_CUDA(cudaMalloc(&g->dh32, R * 4 * C * 4));
  MGV_CHECK_CUDA(cudaMalloc(&g->dlogits, R * static_cast<size_t>(g->Vout) * 4));
  MGV_CHECK_CUDA(cudaMalloc(&g->dtokens, R * static_cast<size_t>(g->cfg.block_size + 1) * 8));
  MGV_CHECK_CUDA(cudaMalloc(&g->dln, R * C * 2));
  MGV_CHECK_CUDA(cudaMalloc(&g->dy, R * C * 2));
  MGV_CHECK_CUDA(cudaMalloc(&g->dh, R * 4 * C * 2));
  const size_t kv_elems = static_cast<size_t>(g->L) * 2 * R * g->nh * g->Tmax * GPT_HEAD_DIM;
  MGV_CHECK_CUDA(cudaMalloc(&g->kv, kv_elems * 2));
  MGV_CHECK_CUDA(cudaMalloc(&g->stats1, static_cast<size_t>(FOLD_MAX_PARTS) * R * sizeof(float2)));
  MGV_CHECK_CUDA(cudaMalloc(&g->stats2, static_cast<size_t>(FOLD_MAX_PARTS) * R * sizeof(float2)));
  g->dec_B = B;
  return MGV_OK;
}

int check_loaded(const Gpt* g) {
  for (int i = 0; i < g->n_tensors; ++i)
    if (!g->loaded[i]) {
      // the class embedder is optional (plain GPT has none)
      if (i == 2) continue;
      set_error("gpt: weight tensor #%d was never loaded (call mgv_gpt_load_weight for every state_dict key)", i);
      return MGV_ERR_STATE;
    }
  return MGV_OK;
}

// layers of the prefill / teacher-forced pass over `rows` = B*T residual rows already in g->x
int run_layers_prefill(Gpt* g, int B, int T, float* att_out, int att_T, bool write_cache, cudaStream_t s) {
  const int C = g->C, rows = B * T;
  for (int l = 0; l < g->L; ++l) {
    const GptLayer& w = g->layers[l];
    MGV_TRY(gpt_layernorm(g->x, w.ln1_w, w.ln1_b, rows, C, g->ln, nullptr, 0, s, false));
    GemmArgs a;
    a.stream = s;
    a.max_stages = 3;   // 96 KB of smem -> two CTAs per SM (epilogue of one overlaps the MMAs of the other)
    a.A = g->ln; a.B = w.wqkv; a.M = rows; a.N = 3 * C; a.K = C;
    a.epi = EPI_BF16; a.bias = w.bqkv; a.out = g->qkv; a.bn = pick_bn(a.N);
    MGV_TRY(gemm_bf16_tc(a));
    float* att = (l == g->L - 1) ? att_out : nullptr;
    MGV_TRY(gpt_attention_prefill(g->qkv, B, T, g->nh, g->cfg.n_unmasked, g->y, att, att_T,
                                  write_cache ? g->kcache(l) : nullptr, write_cache ? g->vcache(l) : nullptr, g->Tmax,
                                  s));
    a = GemmArgs();
    a.stream = s;
    a.max_stages = 3;   // 96 KB of smem -> two CTAs per SM (epilogue of one overlaps the MMAs of the other)
    a.A = g->y; a.B = w.wproj; a.M = rows; a.N = C; a.K = C;
    a.epi = EPI_F32_RESID; a.bias = w.bproj; a.out = g->x; a.resid = g->x; a.bn = pick_bn(a.N);
    MGV_TRY(gemm_bf16_tc(a));
    MGV_TRY(gpt_layernorm(g->x, w.ln2_w, w.ln2_b, rows, C, g->ln, nullptr, 0, s, false));
    a = GemmArgs();
    a.stream = s;
    a.max_stages = 3;   // 96 KB of smem -> two CTAs per SM (epilogue of one overlaps the MMAs of the other)
    a.A = g->ln; a.B = w.wfc1; a.M = rows; a.N = 4 * C; a.K = C;
    a.epi = EPI_BF16_GELU; a.bias = w.bfc1; a.out = g->h; a.bn = pick_bn(a.N);
    MGV_TRY(gemm_bf16_tc(a));
    a = GemmArgs();
    a.stream = s;
    a.max_stages = 3;   // 96 KB of smem -> two CTAs per SM (epilogue of one overlaps the MMAs of the other)
    a.A = g->h; a.B = w.wfc2; a.M = rows; a.N = C; a.K = 4 * C;
    a.epi = EPI_F32_RESID; a.bias = w.bfc2; a.out = g->x; a.resid = g->x; a.bn = pick_bn(a.N);
    MGV_TRY(gemm_bf16_tc(a));
    g->launches += 7;
  }
  return MGV_OK;
}

int read_err_flag(Gpt* g, cudaStream_t s, const char* what) {
  int flag = 0;
  MGV_CHECK_CUDA(cudaMemcpyAsync(&flag, g->d_state + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
  MGV_CHECK_CUDA(cudaStreamSynchronize(s));
  if (flag != 0) {
    MGV_CHECK_CUDA(cudaMemsetAsync(g->d_state + 2, 0, sizeof(int), s));
    set_error("%s: %s index out of range", what, flag == 1 ? "class" : (flag == 2 ? "token / target" : "token"));
    return MGV_ERR_INVALID;
  }
  return MGV_OK;
}

}  // namespace

int gpt_create(const GptConfig* cfg, Gpt** out) {
  MGV_REQUIRE(cfg && out, "gpt_create: null");
  MGV_TRY(check_device());
  MGV_REQUIRE(cfg->n_embd % cfg->n_head == 0 && cfg->n_embd / cfg->n_head == GPT_HEAD_DIM,
              "gpt: head dim %d unsupported (need %d)", cfg->n_head ? cfg->n_embd / cfg->n_head : 0, GPT_HEAD_DIM);
  MGV_REQUIRE(cfg->n_embd % 64 == 0 && cfg->n_embd <= 2048, "gpt: n_embd=%d unsupported", cfg->n_embd);
  MGV_REQUIRE(cfg->block_size >= 1 && cfg->block_size <= GPT_MAX_T, "gpt: block_size=%d unsupported (<= %d)",
              cfg->block_size, GPT_MAX_T);
  MGV_REQUIRE(cfg->vocab_size >= 1 && cfg->n_layer >= 1, "gpt: bad config");
  const int vout = cfg->head_out > 0 ? cfg->head_out : cfg->vocab_size;
  MGV_REQUIRE(vout % 32 == 0, "gpt: head output size %d must be a multiple of 32", vout);
  Gpt* g = new Gpt();
  g->cfg = *cfg;
  g->C = cfg->n_embd;
  g->L = cfg->n_layer;
  g->nh = cfg->n_head;
  g->V = cfg->vocab_size;
  g->Vout = vout;
  g->Tmax = cfg->block_size;
  size_t total = 0;
  carve_params(g, nullptr, &total);
  g->slab_bytes = total;
  if (cudaMalloc(&g->slab, total) != cudaSuccess) {
    set_error("gpt_create: cudaMalloc(%zu) failed", total);
    delete g;
    return MGV_ERR_CUDA;
  }
  carve_params(g, static_cast<char*>(g->slab), &total);
  g->n_tensors = 6 + 12 * g->L;
  g->loaded.assign(g->n_tensors, 0);
  cudaMalloc(&g->d_state, 32 * sizeof(int));
  cudaMemset(g->d_state, 0, 32 * sizeof(int));
  cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&g->ev_in, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&g->ev_out, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming);
  for (int c = 1; c < Gpt::MAX_GROUPS; ++c) {
    cudaStreamCreateWithFlags(&g->gstream[c], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&g->ev_join[c], cudaEventDisableTiming);
  }
  const char* fe = getenv("MGV_DECODE_FOLD");
  if (fe) g->use_fold = atoi(fe) != 0;
  const char* fb = getenv("MGV_FOLD_BN");
  if (fb) {
    int a = 0, b = 0;
    if (sscanf(fb, "%d,%d", &a, &b) == 2 && (a == 32 || a == 64) && (b == 32 || b == 64)) { g->fold_bn = a; g->fold_bn2 = b; }
  }
  const char* fw = getenv("MGV_FOLD_SW");
  if (fw) {
    int a = 0, b = 0;
    if (sscanf(fw, "%d,%d", &a, &b) == 2 && (a == 4 || a == 8) && (b == 4 || b == 8)) { g->fold_sw = a; g->fold_sw_gelu = b; }
  }
  const char* ge = getenv("MGV_DECODE_GROUPS");
  if (ge && atoi(ge) >= 1 && atoi(ge) <= Gpt::MAX_GROUPS) g->groups = atoi(ge);
  const char* e = getenv("MGV_PDL");
  g->pdl = e ? atoi(e) != 0 : true;   // programmatic dependent launch between the decode-step kernels
  const char* tl = getenv("MGV_DECODE_SPLITS");
  if (tl) {
    DecodeTiles t;
    if (sscanf(tl, "%d,%d,%d,%d", &t.qkv_split, &t.proj_split, &t.fc1_split, &t.fc2_split) == 4 && t.qkv_split >= 1 &&
        t.proj_split >= 1 && t.fc1_split >= 1 && t.fc2_split >= 1)
      g->tiles = t;
  }
  if (cudaGetLastError() != cudaSuccess) {
    set_error("gpt_create: CUDA resource creation failed");
    delete g;
    return MGV_ERR_CUDA;
  }
  *out = g;
  return MGV_OK;
}

int gpt_destroy(Gpt* g) {
  if (!g) return MGV_OK;
  gpt_train_release(g);
  cudaFree(g->slab);
  cudaFree(g->x); cudaFree(g->ln); cudaFree(g->qkv); cudaFree(g->y); cudaFree(g->h);
  cudaFree(g->dx); cudaFree(g->dqkv32); cudaFree(g->dh32); cudaFree(g->dln); cudaFree(g->dy); cudaFree(g->dh);
  cudaFree(g->kv); cudaFree(g->dlogits); cudaFree(g->dtokens); cudaFree(g->stats1); cudaFree(g->stats2);
  if (g->graph_exec) cudaGraphExecDestroy(g->graph_exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  cudaFree(g->d_state);
  if (g->stream) cudaStreamDestroy(g->stream);
  if (g->ev_in) cudaEventDestroy(g->ev_in);
  if (g->ev_out) cudaEventDestroy(g->ev_out);
  if (g->ev_fork) cudaEventDestroy(g->ev_fork);
  for (int c = 1; c < Gpt::MAX_GROUPS; ++c) {
    if (g->gstream[c]) cudaStreamDestroy(g->gstream[c]);
    if (g->ev_join[c]) cudaEventDestroy(g->ev_join[c]);
  }
  delete g;
  return MGV_OK;
}

// name = reference state_dict key (transformer/minGPT.py: GPT.__init__ :135-149, Block :97-105,
// CausalSelfAttention :56-63, GPTClass :207).  src = fp32 device pointer.
int gpt_load_weight(Gpt* g, const char* name, const float* src, long long numel, cudaStream_t s) {
  MGV_REQUIRE(g && name && src, "gpt_load_weight: null");
  g->fold_dirty = true;
  const size_t C = g->C;
  auto copy_f32 = [&](float* dst, size_t n, int slot) -> int {
    MGV_REQUIRE(static_cast<size_t>(numel) == n, "gpt_load_weight(%s): numel %lld != expected %zu", name, numel, n);
    MGV_CHECK_CUDA(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, s));
    g->loaded[slot] = 1;
    return MGV_OK;
  };
  auto copy_bf16 = [&](__nv_bfloat16* dst, size_t n, int slot, unsigned char bit) -> int {
    MGV_REQUIRE(static_cast<size_t>(numel) == n, "gpt_load_weight(%s): numel %lld != expected %zu", name, numel, n);
    MGV_TRY(cvt_bf16(src, dst, static_cast<long long>(n), s));
    g->loaded[slot] |= bit;
    return MGV_OK;
  };
  std::string k(name);
  if (k == "tok_emb.weight") return copy_f32(g->tok_emb, static_cast<size_t>(g->V) * C, 0);
  if (k == "pos_emb") return copy_f32(g->pos_emb, static_cast<size_t>(g->cfg.block_size) * C, 1);
  if (k == "embedder.weight") return copy_f32(g->embedder, static_cast<size_t>(g->cfg.class_size) * C, 2);
  if (k == "ln_f.weight") return copy_f32(g->lnf_w, C, 3);
  if (k == "ln_f.bias") return copy_f32(g->lnf_b, C, 4);
  if (k == "head.weight") return copy_bf16(g->whead, static_cast<size_t>(g->Vout) * C, 5, 1);
  int l = -1, consumed = 0;
  if (sscanf(name, "blocks.%d.%n", &l, &consumed) == 1 && consumed > 0 && l >= 0 && l < g->L) {
    GptLayer& y = g->layers[l];
    const std::string r(name + consumed);
    const int base = 6 + 12 * l;
    if (r == "attn.mask") return MGV_OK;  // derived from n_unmasked; not a parameter
    if (r == "ln1.weight") return copy_f32(y.ln1_w, C, base + 0);
    if (r == "ln1.bias") return copy_f32(y.ln1_b, C, base + 1);
    if (r == "ln2.weight") return copy_f32(y.ln2_w, C, base + 2);
    if (r == "ln2.bias") return copy_f32(y.ln2_b, C, base + 3);
    // fused QKV: rows [query | key | value]
    if (r == "attn.query.weight") return copy_bf16(y.wqkv, C * C, base + 4, 1);
    if (r == "attn.key.weight") return copy_bf16(y.wqkv + C * C, C * C, base + 4, 2);
    if (r == "attn.value.weight") return copy_bf16(y.wqkv + 2 * C * C, C * C, base + 4, 4);
    if (r == "attn.query.bias" || r == "attn.key.bias" || r == "attn.value.bias") {
      MGV_REQUIRE(static_cast<size_t>(numel) == C, "gpt_load_weight(%s): numel", name);
      const int part = r[5] == 'q' ? 0 : (r[5] == 'k' ? 1 : 2);
      MGV_CHECK_CUDA(cudaMemcpyAsync(y.bqkv + part * C, src, C * 4, cudaMemcpyDeviceToDevice, s));
      g->loaded[base + 5] |= static_cast<unsigned char>(1 << part);
      return MGV_OK;
    }
    if (r == "attn.proj.weight") return copy_bf16(y.wproj, C * C, base + 6, 1);
    if (r == "attn.proj.bias") return copy_f32(y.bproj, C, base + 7);
    if (r == "mlp.0.weight") return copy_bf16(y.wfc1, 4 * C * C, base + 8, 1);
    if (r == "mlp.0.bias") return copy_f32(y.bfc1, 4 * C, base + 9);
    if (r == "mlp.2.weight") return copy_bf16(y.wfc2, 4 * C * C, base + 10, 1);
    if (r == "mlp.2.bias") return copy_f32(y.bfc2, C, base + 11);
  }
  set_error("gpt_load_weight: unknown tensor name '%s'", name);
  return MGV_ERR_INVALID;
}

namespace {
int check_qkv_complete(const Gpt* g) {
  for (int l = 0; l < g->L; ++l) {
    const int base = 6 + 12 * l;
    if (g->loaded[base + 4] != 7 || g->loaded[base + 5] != 7) {
      set_error("gpt: layer %d query/key/value weights incomplete", l);
      return MGV_ERR_STATE;
    }
  }
  return MGV_OK;
}
}  // namespace

// GPT.forward (minGPT.py:168-199) / GPTClass.forward (:209-212): logits [B, m+t, Vout] fp32,
// att (optional) [B, nh, m+t, m+t] fp32 = last layer's post-softmax attention.
int gpt_forward(Gpt* g, const long long* idx, int B, int t, const float* prefix_emb, const long long* cls, int m,
                float* logits_out, float* att_out, cudaStream_t s) {
  MGV_REQUIRE(g && logits_out, "gpt_forward: null");
  MGV_TRY(check_loaded(g));
  MGV_TRY(check_qkv_complete(g));
  const int T = m + t;
  MGV_REQUIRE(B >= 0 && t >= 0 && m >= 0, "gpt_forward: negative sizes");
  MGV_REQUIRE(T >= 1, "gpt_forward: empty sequence");
  // "Cannot forward, model block size is exhausted." (minGPT.py:178)
  MGV_REQUIRE(T <= g->cfg.block_size, "Cannot forward, model block size is exhausted. (t=%d > block_size=%d)", T,
              g->cfg.block_size);
  MGV_REQUIRE(m == 0 || prefix_emb || (cls && g->loaded[2]), "gpt_forward: prefix without embeddings / embedder");
  if (B == 0) return MGV_OK;
  g->launches = 0;
  const int rows = B * T;
  MGV_TRY(ensure_prefill_ws(g, rows));
  MGV_TRY(gpt_embed(idx, B, T, 0, t, prefix_emb, cls, g->embedder, m, g->tok_emb, g->pos_emb, g->C, g->V,
                    g->cfg.class_size, g->x, g->d_state + 2, s, false));
  if (att_out) MGV_CHECK_CUDA(cudaMemsetAsync(att_out, 0, static_cast<size_t>(B) * g->nh * T * T * 4, s));
  MGV_TRY(run_layers_prefill(g, B, T, att_out, T, false, s));
  MGV_TRY(gpt_layernorm(g->x, g->lnf_w, g->lnf_b, rows, g->C, g->ln, nullptr, 0, s, false));
  GemmArgs a;
  a.stream = s;
  a.A = g->ln; a.B = g->whead; a.M = rows; a.N = g->Vout; a.K = g->C;
  a.epi = EPI_F32; a.bias = nullptr; a.out = logits_out; a.bn = pick_bn(a.N);
  MGV_TRY(gemm_bf16_tc(a));
  g->launches += 3;
  return read_err_flag(g, s, "gpt_forward");
}

namespace {

int decode_gemm(Gpt* g, const __nv_bfloat16* X, const __nv_bfloat16* W, const float* bias, int B, int N, int K,
                int split, int epi_direct, void* out, const void* resid, cudaStream_t s) {
  GemmArgs a;
  a.stream = s;
  a.pdl = g->pdl;
  a.weights_evict_first = true;
  a.bias = bias;
  a.out = out;
  // swap-AB: the weights are the 128-row MMA operand, the B batch rows are the MMA N dimension.  (Measured
  // alternatives, both slower: sequences as the M rows with 16-byte vector reductions, 1144 vs 1016 us per position --
  // a reduction costs per byte, not per instruction; activation tile staged with plain loads instead of TMA, 1025 us.)
  a.transpose_out = true;
  a.A = W; a.B = X; a.M = N; a.N = B; a.K = K;
  // 32 sequences per CTA up to batch 64: twice the CTAs, half the reductions per thread (measured 950 vs 1016 us per
  // position at batch 64; the weight tile is read twice, the second time from L2)
  a.bn = B <= 64 ? 32 : (B <= 128 ? 64 : (B <= 256 ? 128 : 256));
  static const int force_bn = getenv("MGV_DECODE_BN") ? atoi(getenv("MGV_DECODE_BN")) : 0;
  if (force_bn == 32 || force_bn == 64 || force_bn == 128 || force_bn == 256) a.bn = force_bn;
  if (split > 1) {
    a.epi = EPI_F32_ATOMIC;
    a.split_k = split;
  } else {
    a.epi = epi_direct;
    a.resid = resid;
  }
  g->launches += 1;
  return gemm_bf16_tc(a);
}

// one decode position for the B sequences of one group, rows [b0, b0+B) of the batch (enqueued on s; position
// read from sa.pos_ptr)
// K slice (in 64-wide k-blocks) of a fold GEMM CTA: at most 4 resident k-blocks, at most FOLD_MAX_PARTS slices
int fold_kbps(int K) {
  const int nkb = K / 64;
  int kbps = nkb < 4 ? nkb : 4;
  while (ceil_div(nkb, kbps) > FOLD_MAX_PARTS) ++kbps;
  return kbps;
}
int fold_kbps_head(int K) {      // one feature tile only: the finest slices give the most CTAs
  const int nkb = K / 64;
  int kbps = 1;
  while (ceil_div(nkb, kbps) > FOLD_MAX_PARTS) ++kbps;
  return kbps;
}
bool fold_eligible(const Gpt* g, int B) {
  return g->use_fold && B >= 1 && B <= 256 && fold_kbps(4 * g->C) <= 4 && fold_kbps_head(g->C) <= 4;
}

int enqueue_group_step(Gpt* g, int b0, int B, const SampleArgs& sa, float* att_out, int att_T, cudaStream_t s,
                       bool first_kernel_pdl) {
  const int C = g->C;
  const DecodeTiles& tl = g->tiles;
  const size_t r0 = static_cast<size_t>(b0);
  float* dx = g->dx + r0 * C;
  float* dqkv32 = g->dqkv32 + r0 * 3 * C;
  float* dh32 = g->dh32 + r0 * 4 * C;
  float* dlogits = g->dlogits + r0 * g->Vout;
  __nv_bfloat16* dln = g->dln + r0 * C;
  __nv_bfloat16* dy = g->dy + r0 * C;
  __nv_bfloat16* dh = g->dh + r0 * 4 * C;
  const size_t kv_off = r0 * g->nh * g->Tmax * GPT_HEAD_DIM;
  if (att_out) att_out += r0 * g->nh * att_T * att_T;
  if (fold_eligible(g, B)) {
    // ---- folded chain: 5 kernels per block (see gemm_decode_fold.cu)
    const int kq = fold_kbps(C), kf2 = fold_kbps(4 * C);
    const int fbn = (B > 32 && g->fold_bn == 64) ? 64 : 32, fbn2 = (B > 32 && g->fold_bn2 == 64) ? 64 : 32;
    const int np1 = ceil_div(C / 64, kq);                 // K slices of the GEMMs that consume the residual stream
    float2* st1 = g->stats1 + r0;
    float2* st2 = g->stats2 + r0;
    const int sstride = g->dec_B;
    for (int l = 0; l < g->L; ++l) {
      const GptLayer& w = g->layers[l];
      LnFold f1;
      f1.stats = st1; f1.nparts = np1; f1.stride = sstride; f1.dim = C; f1.sw = w.sw_qkv; f1.bp = w.bp_qkv;
      LnFold f2;
      f2.stats = st2; f2.nparts = np1; f2.stride = sstride; f2.dim = C; f2.sw = w.sw_fc1; f2.bp = w.bp_fc1;
      // the first kernel of a position must not start before the previous position's sampler has advanced *pos_ptr
      // (attention reads it ahead of its grid dependency): inside a graph replay positions are serialised anyway
      const bool pdl0 = g->pdl && (l > 0 || first_kernel_pdl);
      MGV_TRY(gemm_decode_fold(FOLD_LN, w.wqkv, 3 * C, C, dx, B, w.ln1_w, st1, sstride, nullptr, nullptr, dqkv32, 3 * C, kq,
                               g->fold_sw, fbn, pdl0, s));
      // (the attention kernel also clears the FC1 accumulator: its last reader, the previous block's FC2, is complete)
      MGV_TRY(gpt_attention_decode(dqkv32, B, g->nh, sa.pos_ptr, g->kcache(l) + kv_off, g->vcache(l) + kv_off, g->Tmax, dy,
                                   (l == g->L - 1) ? att_out : nullptr, att_T, true, dh32,
                                   static_cast<long long>(B) * 4 * C, s, g->pdl, &f1));
      MGV_TRY(decode_gemm(g, dy, w.wproj, w.bproj, B, C, C, tl.proj_split, EPI_F32_RESID, dx, dx, s));
      MGV_TRY(gemm_decode_fold(FOLD_LN, w.wfc1, 4 * C, C, dx, B, w.ln2_w, st2, sstride, nullptr, nullptr, dh32, 4 * C, kq,
                               g->fold_sw, fbn, g->pdl, s));
      MGV_TRY(gemm_decode_fold(FOLD_GELU, w.wfc2, C, 4 * C, dh32, B, nullptr, nullptr, 0, &f2, w.bfc2, dx, C, kf2,
                               g->fold_sw_gelu, fbn2, g->pdl, s));
      g->launches += 4;
    }
    // ln_f + head (minGPT.py:186-188): the same fold; the sampler applies the statistics
    const int kh = fold_kbps_head(C);
    MGV_TRY(gemm_decode_fold(FOLD_LN, g->whead, g->V, C, dx, B, g->lnf_w, st1, sstride, nullptr, nullptr, dlogits, g->V, kh,
                             g->fold_sw, fbn, g->pdl, s));
    SampleArgs sf = sa;
    sf.fold.stats = st1; sf.fold.nparts = ceil_div(C / 64, kh); sf.fold.stride = sstride; sf.fold.dim = C;
    sf.fold.sw = g->sw_head; sf.fold.bp = g->bp_head;
    MGV_TRY(gpt_sample_step(sf, s, g->pdl));
    g->launches += 2;
    return MGV_OK;
  }
  const bool qs = tl.qkv_split > 1, fs = tl.fc1_split > 1;
  for (int l = 0; l < g->L; ++l) {
    const GptLayer& w = g->layers[l];
    MGV_TRY(gpt_layernorm(dx, w.ln1_w, w.ln1_b, B, C, dln, nullptr, 0, s, g->pdl && (l > 0 || first_kernel_pdl)));
    MGV_TRY(decode_gemm(g, dln, w.wqkv, w.bqkv, B, 3 * C, C, tl.qkv_split, EPI_F32, dqkv32, nullptr, s));
    MGV_TRY(gpt_attention_decode(dqkv32, B, g->nh, sa.pos_ptr, g->kcache(l) + kv_off, g->vcache(l) + kv_off, g->Tmax, dy,
                                 (l == g->L - 1) ? att_out : nullptr, att_T, qs, nullptr, 0, s, g->pdl));
    MGV_TRY(decode_gemm(g, dy, w.wproj, w.bproj, B, C, C, tl.proj_split, EPI_F32_RESID, dx, dx, s));
    MGV_TRY(gpt_layernorm(dx, w.ln2_w, w.ln2_b, B, C, dln, nullptr, 0, s, g->pdl));
    static const bool fc1_fullk = getenv("MGV_FC1_FULLK") ? atoi(getenv("MGV_FC1_FULLK")) != 0 : true;
    if (fc1_fullk && C % 256 == 0 && B > 32) {   // at most 32 sequences: 32 CTAs would stream 256 KB each (bs=1: 693 vs 591 us)
      // FC1 without split-K: bias + GELU fused into the epilogue, the separate GELU stage disappears
      MGV_TRY(gemm_decode_fullk(w.wfc1, 4 * C, C, dln, B, w.bfc1, EPI_BF16_GELU, dh, nullptr, 4 * C, g->pdl, s));
      g->launches += 1;
    } else if (fs) {
      MGV_TRY(decode_gemm(g, dln, w.wfc1, w.bfc1, B, 4 * C, C, tl.fc1_split, EPI_F32, dh32, nullptr, s));
      MGV_TRY(gpt_gelu_bf16(dh32, static_cast<long long>(B) * 4 * C, dh, true, s, g->pdl));
      g->launches += 1;
    } else {
      MGV_TRY(decode_gemm(g, dln, w.wfc1, w.bfc1, B, 4 * C, C, 1, EPI_BF16_GELU, dh, nullptr, s));
    }
    MGV_TRY(decode_gemm(g, dh, w.wfc2, w.bfc2, B, C, 4 * C, tl.fc2_split, EPI_F32_RESID, dx, dx, s));
    g->launches += 3;
  }
  // ln_f + head (minGPT.py:186-188) with the same kernels as the blocks, then the fused sampler
  MGV_TRY(gpt_layernorm(dx, g->lnf_w, g->lnf_b, B, C, dln, nullptr, 0, s, g->pdl));
  MGV_TRY(decode_gemm(g, dln, g->whead, nullptr, B, g->V, C, tl.head_split, EPI_F32, dlogits, nullptr, s));
  MGV_TRY(gpt_sample_step(sa, s, g->pdl));
  g->launches += 2;
  return MGV_OK;
}

// fold vectors of every GEMM that consumes a LayerNorm (recomputed after any weight load)
int prepare_fold(Gpt* g, cudaStream_t s) {
  if (!g->use_fold || !g->fold_dirty) return MGV_OK;
  const int C = g->C;
  for (int l = 0; l < g->L; ++l) {
    const GptLayer& w = g->layers[l];
    MGV_TRY(gpt_fold_prepare(w.wqkv, 3 * C, C, w.ln1_w, w.ln1_b, w.bqkv, w.sw_qkv, w.bp_qkv, s));
    MGV_TRY(gpt_fold_prepare(w.wfc1, 4 * C, C, w.ln2_w, w.ln2_b, w.bfc1, w.sw_fc1, w.bp_fc1, s));
  }
  MGV_TRY(gpt_fold_prepare(g->whead, g->Vout, C, g->lnf_w, g->lnf_b, nullptr, g->sw_head, g->bp_head, s));
  g->fold_dirty = false;
  return MGV_OK;
}

int decode_groups(const Gpt* g, int B) {
  int n = g->groups;
  while (n > 1 && B / n < 8) n >>= 1;   // tiny batches: one chain
  return n < 1 ? 1 : n;
}

// one decode position for all B sequences: the sequence groups fork from s, run their chains on their own
// streams and join back (inside a stream capture this becomes parallel branches of the step graph)
int enqueue_decode_step(Gpt* g, int B, const SampleArgs& sa0, float* att_out, int att_T, cudaStream_t s,
                        bool first_kernel_pdl) {
  const int ng = decode_groups(g, B);
  if (ng > 1) MGV_CHECK_CUDA(cudaEventRecord(g->ev_fork, s));
  int b0 = 0;
  for (int c = 0; c < ng; ++c) {
    const int Bc = B / ng + (c < B % ng ? 1 : 0);
    cudaStream_t cs = c == 0 ? s : g->gstream[c];
    if (c > 0) MGV_CHECK_CUDA(cudaStreamWaitEvent(cs, g->ev_fork, 0));
    SampleArgs sa = sa0;
    sa.B = Bc;
    sa.row0 = b0;
    sa.logits_acc = sa0.logits_acc + static_cast<size_t>(b0) * g->V;
    sa.tokens = sa0.tokens + static_cast<size_t>(b0) * sa0.tokens_ld;
    sa.x_next = sa0.x_next + static_cast<size_t>(b0) * g->C;
    sa.pos_ptr = g->d_state + 8 + 2 * c;
    sa.done_counter = reinterpret_cast<unsigned int*>(g->d_state + 9 + 2 * c);
    const int rc = enqueue_group_step(g, b0, Bc, sa, att_out, att_T, cs, first_kernel_pdl);
    if (rc != MGV_OK) return rc;
    if (c > 0) MGV_CHECK_CUDA(cudaEventRecord(g->ev_join[c], cs));
    b0 += Bc;
  }
  for (int c = 1; c < ng; ++c) MGV_CHECK_CUDA(cudaStreamWaitEvent(s, g->ev_join[c], 0));
  return MGV_OK;
}

}  // namespace

// Lit_minGPT.sample (minGPT.py:293-360) with a KV cache: x_out [B, t0+steps] int64 (first t0
// columns = x0), att_out (optional) [B, nh, Tf, Tf] fp32 with Tf = m + t0 + steps - 1 = the
// sequence length of the reference's last forward call.
int gpt_generate(Gpt* g, const long long* x0, int B, int t0, const float* prefix_emb, const long long* cls, int m,
                 int steps, float temperature, int do_sample, int top_k, unsigned long long seed, long long* x_out,
                 float* att_out, int use_graph, cudaStream_t caller) {
  MGV_REQUIRE(g && x_out, "gpt_generate: null");
  MGV_TRY(check_loaded(g));
  MGV_TRY(check_qkv_complete(g));
  MGV_REQUIRE(B >= 0 && t0 >= 0 && steps >= 0, "gpt_generate: negative sizes");
  MGV_REQUIRE(m >= 0 && m + t0 >= 1, "gpt_generate: empty context (needs a conditioning prefix or a prompt)");
  MGV_REQUIRE(m == 0 || prefix_emb || (cls && g->loaded[2]), "gpt_generate: prefix without embeddings / embedder");
  MGV_REQUIRE(g->Vout == g->V, "gpt_generate: head output %d != vocab %d", g->Vout, g->V);
  MGV_REQUIRE(top_k >= 0 && top_k <= g->V, "gpt_generate: top_k=%d out of range for vocab %d", top_k, g->V);
  // assert x.size(1) + cond_size <= block_size at every step (minGPT.py:336)
  MGV_REQUIRE(steps == 0 || t0 + steps - 1 + m <= g->cfg.block_size,
              "sample: context %d + cond %d exceeds block_size %d", t0 + steps - 1, m, g->cfg.block_size);
  // KV-cache decoding equals the reference's per-step full forward only for a causal mask: with an unmasked prefix
  // (minGPT.py:67-68) rows below n_unmasked would re-attend to every later token at every step
  MGV_REQUIRE(g->cfg.n_unmasked <= 1, "gpt_generate: n_unmasked=%d > 1 is not supported by the KV-cache decode loop "
              "(the reference recomputes the prefix-unmasked rows at every step)", g->cfg.n_unmasked);
  if (B == 0) return MGV_OK;
  g->launches = 0;
  cudaStream_t s = g->stream;
  MGV_CHECK_CUDA(cudaEventRecord(g->ev_in, caller));
  MGV_CHECK_CUDA(cudaStreamWaitEvent(s, g->ev_in, 0));
  MGV_TRY(ensure_decode_ws(g, B));
  const int ld = t0 + steps;                 // row length of the caller's x_out
  const int tld = g->cfg.block_size + 1;     // row length of the internal token buffer
  if (t0 > 0) {
    MGV_CHECK_CUDA(cudaMemcpy2DAsync(x_out, static_cast<size_t>(ld) * 8, x0, static_cast<size_t>(t0) * 8,
                                     static_cast<size_t>(t0) * 8, B, cudaMemcpyDeviceToDevice, s));
    MGV_CHECK_CUDA(cudaMemcpy2DAsync(g->dtokens, static_cast<size_t>(tld) * 8, x0, static_cast<size_t>(t0) * 8,
                                     static_cast<size_t>(t0) * 8, B, cudaMemcpyDeviceToDevice, s));
  }
  if (steps == 0) {
    MGV_CHECK_CUDA(cudaEventRecord(g->ev_out, s));
    MGV_CHECK_CUDA(cudaStreamWaitEvent(caller, g->ev_out, 0));
    return MGV_OK;
  }
  const int T0 = m + t0;              // rows known before sampling starts
  const int Tf = T0 + steps - 1;      // rows of the reference's final forward
  if (att_out) MGV_CHECK_CUDA(cudaMemsetAsync(att_out, 0, static_cast<size_t>(B) * g->nh * Tf * Tf * 4, s));
  // (the KV cache is laid out with dec_B as the batch extent; a smaller B uses a prefix of it)
  // ---- prefill rows [0, T0-1): fills the KV cache only
  if (T0 - 1 > 0) {
    const int Tp = T0 - 1;
    MGV_TRY(ensure_prefill_ws(g, B * Tp));
    MGV_TRY(gpt_embed(x0, B, Tp, 0, t0, prefix_emb, cls, g->embedder, m, g->tok_emb, g->pos_emb, g->C, g->V,
                      g->cfg.class_size, g->x, g->d_state + 2, s, false));
    MGV_TRY(run_layers_prefill(g, B, Tp, att_out, Tf, true, s));
    g->launches += 1;
  }
  // ---- embedding of row T0-1 -> decode residual stream; pos = T0-1
  MGV_TRY(gpt_embed(x0, B, 1, T0 - 1, t0, prefix_emb, cls, g->embedder, m, g->tok_emb, g->pos_emb, g->C, g->V,
                    g->cfg.class_size, g->dx, g->d_state + 2, s, false));
  g->launches += 1;
  // split-K accumulators start at zero; afterwards their consumers clear them for the next layer / position
  MGV_CHECK_CUDA(cudaMemsetAsync(g->dqkv32, 0, static_cast<size_t>(B) * 3 * g->C * 4, s));
  MGV_CHECK_CUDA(cudaMemsetAsync(g->dh32, 0, static_cast<size_t>(B) * 4 * g->C * 4, s));
  MGV_CHECK_CUDA(cudaMemsetAsync(g->dlogits, 0, static_cast<size_t>(B) * g->Vout * 4, s));
  int init_state[2 * Gpt::MAX_GROUPS];
  for (int c = 0; c < Gpt::MAX_GROUPS; ++c) { init_state[2 * c] = T0 - 1; init_state[2 * c + 1] = 0; }
  MGV_CHECK_CUDA(cudaMemcpyAsync(g->d_state + 8, init_state, sizeof(init_state), cudaMemcpyHostToDevice, s));
  MGV_CHECK_CUDA(cudaMemcpyAsync(g->d_state + 4, &seed, sizeof(seed), cudaMemcpyHostToDevice, s));

  SampleArgs sa{};
  sa.logits_acc = g->dlogits;
  sa.B = B; sa.C = g->C; sa.V = g->V;
  sa.temperature = temperature; sa.top_k = top_k; sa.do_sample = do_sample;
  sa.seed_ptr = reinterpret_cast<const unsigned long long*>(g->d_state + 4);
  sa.pos_ptr = g->d_state + 8;
  sa.tokens = g->dtokens; sa.tokens_ld = tld; sa.m = m;
  sa.tok_emb = g->tok_emb; sa.pos_emb = g->pos_emb; sa.block_size = g->cfg.block_size;
  sa.x_next = g->dx;
  sa.logits_out = g->step_logits; sa.logits_step_stride = static_cast<long long>(B) * g->V; sa.logits_pos0 = T0 - 1;
  const bool want_step_logits = g->step_logits != nullptr;
  g->step_logits = nullptr;
  sa.done_counter = reinterpret_cast<unsigned int*>(g->d_state + 9);

  MGV_TRY(prepare_fold(g, s));
  cudaGraph_t tmp_graph = nullptr;
  cudaGraphExec_t tmp_exec = nullptr;
  if (use_graph) {
    // the decode-step graph only depends on (B, m, sampler settings) and on handle-owned buffers, so it is
    // captured once and replayed for every position of every later call with the same settings; a request for
    // the attention map (caller-owned buffer) gets a one-off graph
    const bool cacheable = att_out == nullptr && !want_step_logits;
    const bool hit = cacheable && g->graph_exec && g->graph_key.B == B && g->graph_key.m == m &&
                     g->graph_key.top_k == top_k && g->graph_key.do_sample == do_sample &&
                     g->graph_key.temperature == temperature;
    cudaGraphExec_t exec = hit ? g->graph_exec : nullptr;
    long long per_step = hit ? g->graph_key.per_step : 0;
    if (!hit) {
      cudaGraph_t graph = nullptr;
      MGV_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      const long long before = g->launches;
      const int rc = enqueue_decode_step(g, B, sa, att_out, Tf, s, true);
      per_step = g->launches - before;
      g->launches = before;
      const cudaError_t ce = cudaStreamEndCapture(s, &graph);
      if (rc != MGV_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      MGV_CHECK_CUDA(ce);
      MGV_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
      if (cacheable) {
        if (g->graph_exec) cudaGraphExecDestroy(g->graph_exec);
        if (g->graph) cudaGraphDestroy(g->graph);
        g->graph = graph;
        g->graph_exec = exec;
        g->graph_key = {B, m, top_k, do_sample, temperature, per_step};
      } else {
        tmp_graph = graph;
        tmp_exec = exec;
      }
    }
    for (int k = 0; k < steps; ++k) {
      const cudaError_t le = cudaGraphLaunch(exec, s);
      if (le != cudaSuccess) {
        if (tmp_exec) cudaGraphExecDestroy(tmp_exec);
        if (tmp_graph) cudaGraphDestroy(tmp_graph);
        MGV_CHECK_CUDA(le);
      }
    }
    g->launches += per_step * steps;
  } else {
    for (int k = 0; k < steps; ++k) MGV_TRY(enqueue_decode_step(g, B, sa, att_out, Tf, s, false));
  }
  // generated tokens: internal buffer columns [t0, t0+steps) -> x_out
  const cudaError_t e0 = cudaMemcpy2DAsync(x_out + t0, static_cast<size_t>(ld) * 8, g->dtokens + t0, static_cast<size_t>(tld) * 8,
                                           static_cast<size_t>(steps) * 8, B, cudaMemcpyDeviceToDevice, s);
  const cudaError_t e1 = cudaEventRecord(g->ev_out, s);
  const cudaError_t e2 = cudaStreamWaitEvent(caller, g->ev_out, 0);
  const int rc = read_err_flag(g, s, "gpt_generate");  // synchronises s
  if (tmp_exec) cudaGraphExecDestroy(tmp_exec);
  if (tmp_graph) cudaGraphDestroy(tmp_graph);
  MGV_CHECK_CUDA(e0);
  MGV_CHECK_CUDA(e1);
  MGV_CHECK_CUDA(e2);
  return rc;
}

// per-row cross entropy of fp32 logits [rows, V] against int64 targets (minGPT.py:197, decoders.py:64-68)
int gpt_cross_entropy(Gpt* g, const float* logits, const long long* targets, long long rows, int V, float* loss,
                      cudaStream_t s) {
  MGV_REQUIRE(g, "gpt_cross_entropy: null handle");
  MGV_REQUIRE(rows >= 0, "gpt_cross_entropy: negative rows");
  if (rows == 0) return MGV_OK;
  MGV_TRY(gpt_ce_rows(logits, targets, rows, V, V, loss, g->d_state + 2, s));
  g->launches = 1;
  return read_err_flag(g, s, "gpt_cross_entropy");
}

long long gpt_last_launches(const Gpt* g) { return g ? g->launches : 0; }

// Deterministic decode: every decode GEMM owns its full K (no split-K reductions, whose fp32 summation order depends on
// CTA timing) through the separate-LayerNorm chain; same results run to run and independent of the batch size.
int gpt_set_deterministic(Gpt* g, int on) {
  MGV_REQUIRE(g, "gpt_set_deterministic: null handle");
  const bool want = on != 0;
  if (want == g->deterministic) return MGV_OK;
  g->deterministic = want;
  if (want) {
    g->saved_fold = g->use_fold;
    g->saved_tiles = g->tiles;
    g->saved_groups = g->groups;
    g->use_fold = false;
    g->groups = 1;
    g->tiles.qkv_split = g->tiles.proj_split = g->tiles.fc1_split = g->tiles.fc2_split = g->tiles.head_split = 1;
  } else {
    g->use_fold = g->saved_fold;
    g->tiles = g->saved_tiles;
    g->groups = g->saved_groups;
  }
  if (g->graph_exec) { cudaGraphExecDestroy(g->graph_exec); g->graph_exec = nullptr; }
  if (g->graph) { cudaGraphDestroy(g->graph); g->graph = nullptr; }
  return MGV_OK;
}

int gpt_set_step_logits(Gpt* g, float* buf) {
  MGV_REQUIRE(g, "gpt_set_step_logits: null handle");
  g->step_logits = buf;
  return MGV_OK;
}

}  // namespace mgv
