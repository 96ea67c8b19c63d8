// minGPT training step on the libmgv handle: flat fp32 master parameters / gradients owned by the caller, saved
// activations and transposed weight copies owned by the handle.  See gpt_train.cuh for the GEMM formulation.
// reference: Lit_minGPT.training_step / shared_step transformer/minGPT.py:413-422, GPT.forward :168-199,
// configure_optimizers :618-665 (AdamW, decay / no-decay groups), DDP gradient averaging GPT_VAE_train.py:172-174.
#include <string>
#include <vector>
#include "gemm_tc.cuh"
#include "gpt_impl.cuh"
#include "gpt_train.cuh"

namespace mgv {

struct TrainLayerActs {
  float *x0, *x1;                         // residual stream entering the block / after the attention branch  [R, C]
  __nv_bfloat16 *a, *qkv, *y, *c, *hpre, *h;   // ln1 out, q|k|v, attention out, ln2 out, FC1 out, gelu(FC1 out)
  float* lse2;                            // [B*nh, T]
};

struct GptTrain {
  float* params = nullptr;                // flat fp32 masters (caller-owned, layout = seg table)
  float* grads = nullptr;                 // flat fp32 gradients (caller-owned)
  std::vector<AdamSeg> segs;
  std::vector<std::string> names;
  long long total = 0;
  AdamSeg* d_segs = nullptr;
  int2* d_chunks = nullptr;
  int n_chunks = 0;
  // transposed bf16 weights for the dgrad GEMMs
  void* tslab = nullptr;
  std::vector<__nv_bfloat16*> wqkvT, wprojT, wfc1T, wfc2T;
  __nv_bfloat16* wheadT = nullptr;
  // activations
  void* aslab = nullptr;
  size_t aslab_bytes = 0;
  int cap_B = 0, cap_T = 0;
  std::vector<TrainLayerActs> acts;
  float *xf = nullptr, *logits = nullptr, *tmp32 = nullptr, *dx = nullptr, *delta = nullptr, *loss = nullptr;
  __nv_bfloat16 *f = nullptr, *dlogits = nullptr, *g = nullptr, *gT = nullptr, *actT = nullptr, *wide = nullptr, *dqkv = nullptr;
  // last forward
  int B = 0, T = 0, t = 0, m = 0, R = 0, Rpad = 0;
  const long long* idx = nullptr;
  const long long* cls = nullptr;
  DropCfg d_embd, d_resid, d_attn;
  bool have_forward = false;
};

namespace {

size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

// canonical flat layout: embeddings, then per block (LayerNorms, q|k|v weights adjacent so that the fused QKV
// weight gradient [3C, C] lands in one piece, biases likewise), then ln_f, head.
void build_segments(Gpt* g, GptTrain* tr) {
  const long long C = g->C;
  long long off = 0;
  auto add = [&](const std::string& name, long long numel, int decay, float* f32, __nv_bfloat16* b16) {
    AdamSeg s;
    s.offset = off; s.numel = numel; s.decay = decay; s.dst_f32 = f32; s.dst_bf16 = b16;
    tr->segs.push_back(s);
    tr->names.push_back(name);
    off += (numel + 63) & ~63LL;          // 256-byte aligned segments
  };
  add("tok_emb.weight", static_cast<long long>(g->V) * C, 0, g->tok_emb, nullptr);
  add("pos_emb", static_cast<long long>(g->cfg.block_size) * C, 0, g->pos_emb, nullptr);
  if (g->cfg.class_size > 0) add("embedder.weight", static_cast<long long>(g->cfg.class_size) * C, 0, g->embedder, nullptr);
  for (int l = 0; l < g->L; ++l) {
    GptLayer& y = g->layers[l];
    const std::string p = "blocks." + std::to_string(l) + ".";
    add(p + "ln1.weight", C, 0, y.ln1_w, nullptr);
    add(p + "ln1.bias", C, 0, y.ln1_b, nullptr);
    add(p + "attn.query.weight", C * C, 1, nullptr, y.wqkv);
    add(p + "attn.key.weight", C * C, 1, nullptr, y.wqkv + C * C);
    add(p + "attn.value.weight", C * C, 1, nullptr, y.wqkv + 2 * C * C);
    add(p + "attn.query.bias", C, 0, y.bqkv, nullptr);
    add(p + "attn.key.bias", C, 0, y.bqkv + C, nullptr);
    add(p + "attn.value.bias", C, 0, y.bqkv + 2 * C, nullptr);
    add(p + "attn.proj.weight", C * C, 1, nullptr, y.wproj);
    add(p + "attn.proj.bias", C, 0, y.bproj, nullptr);
    add(p + "ln2.weight", C, 0, y.ln2_w, nullptr);
    add(p + "ln2.bias", C, 0, y.ln2_b, nullptr);
    add(p + "mlp.0.weight", 4 * C * C, 1, nullptr, y.wfc1);
    add(p + "mlp.0.bias", 4 * C, 0, y.bfc1, nullptr);
    add(p + "mlp.2.weight", 4 * C * C, 1, nullptr, y.wfc2);
    add(p + "mlp.2.bias", C, 0, y.bfc2, nullptr);
  }
  add("ln_f.weight", C, 0, g->lnf_w, nullptr);
  add("ln_f.bias", C, 0, g->lnf_b, nullptr);
  add("head.weight", static_cast<long long>(g->Vout) * C, 1, nullptr, g->whead);
  tr->total = off;
}

const AdamSeg* find_seg(const GptTrain* tr, const std::string& name) {
  for (size_t i = 0; i < tr->names.size(); ++i)
    if (tr->names[i] == name) return &tr->segs[i];
  return nullptr;
}
float* grad_of(const GptTrain* tr, const std::string& name) { return tr->grads + find_seg(tr, name)->offset; }

GptTrain* ensure_train(Gpt* g) {
  if (g->train == nullptr) {
    g->train = new GptTrain();
    build_segments(g, g->train);
  }
  return g->train;
}

int alloc_transposed(Gpt* g, GptTrain* tr) {
  if (tr->tslab) return MGV_OK;
  const size_t C = g->C;
  size_t total = 0;
  auto sz = [&](size_t n) { const size_t o = total; total += al256(n * 2); return o; };
  std::vector<size_t> o1(g->L), o2(g->L), o3(g->L), o4(g->L);
  for (int l = 0; l < g->L; ++l) { o1[l] = sz(3 * C * C); o2[l] = sz(C * C); o3[l] = sz(4 * C * C); o4[l] = sz(4 * C * C); }
  const size_t oh = sz(static_cast<size_t>(g->Vout) * C);
  MGV_CHECK_CUDA(cudaMalloc(&tr->tslab, total));
  char* b = static_cast<char*>(tr->tslab);
  tr->wqkvT.resize(g->L); tr->wprojT.resize(g->L); tr->wfc1T.resize(g->L); tr->wfc2T.resize(g->L);
  for (int l = 0; l < g->L; ++l) {
    tr->wqkvT[l] = reinterpret_cast<__nv_bfloat16*>(b + o1[l]);
    tr->wprojT[l] = reinterpret_cast<__nv_bfloat16*>(b + o2[l]);
    tr->wfc1T[l] = reinterpret_cast<__nv_bfloat16*>(b + o3[l]);
    tr->wfc2T[l] = reinterpret_cast<__nv_bfloat16*>(b + o4[l]);
  }
  tr->wheadT = reinterpret_cast<__nv_bfloat16*>(b + oh);
  return MGV_OK;
}

// W^T copies from the handle's bf16 weights (after load_weight or an optimizer step)
int refresh_transposed(Gpt* g, GptTrain* tr, cudaStream_t s) {
  MGV_TRY(alloc_transposed(g, tr));
  const int C = g->C;
  for (int l = 0; l < g->L; ++l) {
    const GptLayer& w = g->layers[l];
    MGV_TRY(train_transpose(w.wqkv, 3 * C, C, C, 3 * C, tr->wqkvT[l], s));       // [3C, C] -> [C, 3C]
    MGV_TRY(train_transpose(w.wproj, C, C, C, C, tr->wprojT[l], s));
    MGV_TRY(train_transpose(w.wfc1, 4 * C, C, C, 4 * C, tr->wfc1T[l], s));       // [4C, C] -> [C, 4C]
    MGV_TRY(train_transpose(w.wfc2, C, 4 * C, 4 * C, C, tr->wfc2T[l], s));       // [C, 4C] -> [4C, C]
  }
  MGV_TRY(train_transpose(g->whead, g->Vout, C, C, g->Vout, tr->wheadT, s));     // [V, C] -> [C, V]
  return MGV_OK;
}

int ensure_acts(Gpt* g, GptTrain* tr, int B, int T) {
  if (B <= tr->cap_B && T <= tr->cap_T && tr->aslab) return MGV_OK;
  if (tr->aslab) { cudaFree(tr->aslab); tr->aslab = nullptr; }
  const int cB = B > tr->cap_B ? B : tr->cap_B, cT = T > tr->cap_T ? T : tr->cap_T;
  const size_t R = static_cast<size_t>(cB) * cT, C = g->C, Rp = (R + 63) & ~size_t(63), V = g->Vout, nh = g->nh;
  for (int pass = 0; pass < 2; ++pass) {
    size_t off = 0;
    char* base = static_cast<char*>(tr->aslab);
    auto take = [&](size_t bytes) -> void* {
      void* p = base ? base + off : nullptr;
      off += al256(bytes);
      return p;
    };
    tr->acts.resize(g->L);
    for (int l = 0; l < g->L; ++l) {
      TrainLayerActs& a = tr->acts[l];
      a.x0 = static_cast<float*>(take(R * C * 4));
      a.x1 = static_cast<float*>(take(R * C * 4));
      a.a = static_cast<__nv_bfloat16*>(take(R * C * 2));
      a.qkv = static_cast<__nv_bfloat16*>(take(R * 3 * C * 2));
      a.y = static_cast<__nv_bfloat16*>(take(R * C * 2));
      a.c = static_cast<__nv_bfloat16*>(take(R * C * 2));
      a.hpre = static_cast<__nv_bfloat16*>(take(R * 4 * C * 2));
      a.h = static_cast<__nv_bfloat16*>(take(R * 4 * C * 2));
      a.lse2 = static_cast<float*>(take(static_cast<size_t>(cB) * nh * cT * 4));
    }
    tr->xf = static_cast<float*>(take(R * C * 4));
    tr->f = static_cast<__nv_bfloat16*>(take(R * C * 2));
    tr->logits = static_cast<float*>(take(R * V * 4));
    tr->dlogits = static_cast<__nv_bfloat16*>(take(R * V * 2));
    tr->tmp32 = static_cast<float*>(take(R * C * 4));
    tr->dx = static_cast<float*>(take(R * C * 4));
    tr->delta = static_cast<float*>(take(static_cast<size_t>(cB) * nh * cT * 4));
    tr->loss = static_cast<float*>(take(256));
    tr->g = static_cast<__nv_bfloat16*>(take(R * 4 * C * 2));          // row-major gradient operand (up to 4C wide)
    tr->gT = static_cast<__nv_bfloat16*>(take(4 * C * Rp * 2));        // transposed gradient operand
    tr->actT = static_cast<__nv_bfloat16*>(take(4 * C * Rp * 2));      // transposed activation operand
    tr->wide = static_cast<__nv_bfloat16*>(take(R * 4 * C * 2));       // dh
    tr->dqkv = static_cast<__nv_bfloat16*>(take(R * 3 * C * 2));
    if (pass == 0) {
      tr->aslab_bytes = off;
      if (cudaMalloc(&tr->aslab, off) != cudaSuccess) {
        tr->aslab = nullptr;
        set_error("gpt train: cudaMalloc(%zu bytes of activations for B=%d, T=%d) failed", off, cB, cT);
        return MGV_ERR_CUDA;
      }
    }
  }
  tr->cap_B = cB;
  tr->cap_T = cT;
  return MGV_OK;
}

int gemm(const void* A, const void* Bm, int M, int N, int K, int epi, const float* bias, void* out, const void* resid,
         cudaStream_t s, long long lda = 0) {
  GemmArgs a;
  a.stream = s;
  a.max_stages = 3;
  a.A = A; a.B = Bm; a.M = M; a.N = N; a.K = K; a.lda = lda;
  a.epi = epi; a.bias = bias; a.out = out; a.resid = resid;
  // tile width: the GEMMs of a per-GPU batch of 8 clips give only 30-270 tiles of 128 x 256 on 148 SMs, so pick the width
  // that minimises (rounds over the SMs) x (time of one tile).  A tile's k-block is bound by the L2 -> shared-memory operand
  // feed (128 + bn rows of 128 bytes), not by its MMAs, hence the (128 + bn) weight; ties go to the wider tile.
  const int n_sm = num_sms();
  const int m_tiles = ceil_div(M, 128);
  int best_bn = 0;
  long long best_cost = 0;
  for (int bn : {256, 128, 64}) {
    if (N % bn != 0) continue;
    const long long tiles = static_cast<long long>(m_tiles) * (N / bn);
    const long long cost = ((tiles + n_sm - 1) / n_sm) * (128 + bn);
    if (best_bn == 0 || cost < best_cost) { best_bn = bn; best_cost = cost; }
  }
  a.bn = best_bn ? best_bn : 32;
  return gemm_bf16_tc(a);
}

}  // namespace

void gpt_train_release(Gpt* g) {
  GptTrain* tr = g->train;
  if (!tr) return;
  cudaFree(tr->tslab);
  cudaFree(tr->aslab);
  cudaFree(tr->d_segs);
  cudaFree(tr->d_chunks);
  delete tr;
  g->train = nullptr;
}

long long gpt_train_numel(Gpt* g) { return g ? ensure_train(g)->total : 0; }

int gpt_train_layout(Gpt* g, const char* name, long long* offset, long long* numel, int* decay) {
  MGV_REQUIRE(g && name && offset && numel, "gpt_train_layout: null");
  const AdamSeg* s = find_seg(ensure_train(g), name);
  if (!s) {
    set_error("gpt_train_layout: unknown parameter '%s'", name);
    return MGV_ERR_INVALID;
  }
  *offset = s->offset;
  *numel = s->numel;
  if (decay) *decay = s->decay;
  return MGV_OK;
}

// flat_params / flat_grads: fp32 device buffers of gpt_train_numel() elements owned by the caller.  The handle's inference
// copies must already hold the same values (mgv_gpt_load_weight of every tensor); from here on the optimizer keeps them in sync.
int gpt_train_bind(Gpt* g, float* flat_params, float* flat_grads, cudaStream_t s) {
  MGV_REQUIRE(g && flat_params && flat_grads, "gpt_train_bind: null");
  GptTrain* tr = ensure_train(g);
  tr->params = flat_params;
  tr->grads = flat_grads;
  if (!tr->d_segs) {
    std::vector<int2> chunks;
    for (size_t i = 0; i < tr->segs.size(); ++i)
      for (long long c = 0; c * ADAM_CHUNK < tr->segs[i].numel; ++c) chunks.push_back(make_int2(static_cast<int>(i), static_cast<int>(c)));
    tr->n_chunks = static_cast<int>(chunks.size());
    MGV_CHECK_CUDA(cudaMalloc(&tr->d_segs, tr->segs.size() * sizeof(AdamSeg)));
    MGV_CHECK_CUDA(cudaMalloc(&tr->d_chunks, chunks.size() * sizeof(int2)));
    MGV_CHECK_CUDA(cudaMemcpyAsync(tr->d_segs, tr->segs.data(), tr->segs.size() * sizeof(AdamSeg), cudaMemcpyHostToDevice, s));
    MGV_CHECK_CUDA(cudaMemcpyAsync(tr->d_chunks, chunks.data(), chunks.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
    MGV_CHECK_CUDA(cudaStreamSynchronize(s));
  }
  return refresh_transposed(g, tr, s);
}

// Teacher-forced forward with dropout; saves what the backward needs; *loss_out (device) = mean cross entropy.
// idx [B, t] int64, cls [B] int64 (m = 1) or null (m = 0), targets [B, m + t] int64.
int gpt_train_forward(Gpt* g, const long long* idx, int B, int t, const long long* cls, int m, const long long* targets,
                      float p_embd, float p_resid, float p_attn, unsigned long long seed, float* loss_out, cudaStream_t s) {
  MGV_REQUIRE(g && g->train && g->train->params, "gpt_train_forward: call mgv_gpt_train_bind first");
  MGV_REQUIRE(idx && targets && loss_out && B >= 1 && t >= 1 && (m == 0 || m == 1), "gpt_train_forward: bad arguments");
  MGV_REQUIRE(m == 0 || (cls && g->cfg.class_size > 0), "gpt_train_forward: class prefix without an embedder");
  MGV_REQUIRE(p_embd >= 0.f && p_embd < 1.f && p_resid >= 0.f && p_resid < 1.f && p_attn >= 0.f && p_attn < 1.f,
              "gpt_train_forward: dropout probabilities must lie in [0, 1)");
  const int T = m + t;
  MGV_REQUIRE(T <= g->cfg.block_size, "Cannot forward, model block size is exhausted. (t=%d > block_size=%d)", T, g->cfg.block_size);
  MGV_REQUIRE(g->Vout % 64 == 0, "gpt_train_forward: head width %d must be a multiple of 64", g->Vout);
  GptTrain* tr = g->train;
  MGV_TRY(ensure_acts(g, tr, B, T));
  const int C = g->C, R = B * T, V = g->Vout;
  tr->B = B; tr->T = T; tr->t = t; tr->m = m; tr->R = R; tr->Rpad = (R + 63) & ~63;
  tr->idx = idx; tr->cls = cls;
  tr->d_embd = make_drop(p_embd, seed);
  tr->d_resid = make_drop(p_resid, seed);
  tr->d_attn = make_drop(p_attn, seed);
  g->launches = 0;
  MGV_CHECK_CUDA(cudaMemsetAsync(tr->loss, 0, sizeof(float), s));
  MGV_TRY(train_embed(idx, B, T, t, cls, g->embedder, m, g->tok_emb, g->pos_emb, C, g->V, g->cfg.class_size, tr->acts[0].x0,
                      g->d_state + 2, tr->d_embd, s));
  const bool rdrop = tr->d_resid.thresh24 != 0u;
  for (int l = 0; l < g->L; ++l) {
    const GptLayer& w = g->layers[l];
    TrainLayerActs& a = tr->acts[l];
    float* x2 = (l + 1 < g->L) ? tr->acts[l + 1].x0 : tr->xf;
    MGV_TRY(gpt_layernorm(a.x0, w.ln1_w, w.ln1_b, R, C, a.a, nullptr, 0, s, false));
    MGV_TRY(gemm(a.a, w.wqkv, R, 3 * C, C, EPI_BF16, w.bqkv, a.qkv, nullptr, s));
    MGV_TRY(train_attn_fwd(a.qkv, B, T, g->nh, g->cfg.n_unmasked, a.y, a.lse2, tr->d_attn, drop_stream_attn(l), s));
    if (rdrop) {
      MGV_TRY(gemm(a.y, w.wproj, R, C, C, EPI_F32, w.bproj, tr->tmp32, nullptr, s));
      MGV_TRY(train_resid_dropout(a.x0, tr->tmp32, static_cast<long long>(R) * C, a.x1, tr->d_resid, drop_stream_resid_attn(l), s));
    } else {
      MGV_TRY(gemm(a.y, w.wproj, R, C, C, EPI_F32_RESID, w.bproj, a.x1, a.x0, s));
    }
    MGV_TRY(gpt_layernorm(a.x1, w.ln2_w, w.ln2_b, R, C, a.c, nullptr, 0, s, false));
    MGV_TRY(gemm(a.c, w.wfc1, R, 4 * C, C, EPI_BF16, w.bfc1, a.hpre, nullptr, s));
    MGV_TRY(train_gelu_fwd(a.hpre, static_cast<long long>(R) * 4 * C, a.h, s));
    if (rdrop) {
      MGV_TRY(gemm(a.h, w.wfc2, R, C, 4 * C, EPI_F32, w.bfc2, tr->tmp32, nullptr, s));
      MGV_TRY(train_resid_dropout(a.x1, tr->tmp32, static_cast<long long>(R) * C, x2, tr->d_resid, drop_stream_resid_mlp(l), s));
    } else {
      MGV_TRY(gemm(a.h, w.wfc2, R, C, 4 * C, EPI_F32_RESID, w.bfc2, x2, a.x1, s));
    }
    g->launches += rdrop ? 10 : 8;
  }
  MGV_TRY(gpt_layernorm(tr->xf, g->lnf_w, g->lnf_b, R, C, tr->f, nullptr, 0, s, false));
  MGV_TRY(gemm(tr->f, g->whead, R, V, C, EPI_F32, nullptr, tr->logits, nullptr, s));
  MGV_TRY(train_ce(tr->logits, targets, R, V, tr->loss, tr->dlogits, g->d_state + 2, s));
  MGV_CHECK_CUDA(cudaMemcpyAsync(loss_out, tr->loss, sizeof(float), cudaMemcpyDeviceToDevice, s));
  g->launches += 4;
  tr->have_forward = true;
  return MGV_OK;
}

// Backward of the last forward for blocks layer_hi-1 ... layer_lo (layer_hi == n_layer also runs the head / ln_f part and
// clears the gradient buffer first; layer_lo == 0 also runs the embedding backward).  Splitting the call lets the host launch
// the all-reduce of a finished bucket of layers (the flat gradient layout is layer-contiguous) while the next bucket runs.
int gpt_train_backward(Gpt* g, int layer_hi, int layer_lo, cudaStream_t s) {
  MGV_REQUIRE(g && g->train && g->train->have_forward, "gpt_train_backward: no forward to differentiate");
  GptTrain* tr = g->train;
  MGV_REQUIRE(layer_lo >= 0 && layer_lo <= layer_hi && layer_hi <= g->L, "gpt_train_backward: bad layer range [%d, %d)", layer_lo, layer_hi);
  const int C = g->C, R = tr->R, Rp = tr->Rpad, V = g->Vout, B = tr->B, T = tr->T;
  const DropCfg none = make_drop(0.f, 0);
  auto gname = [&](int l, const char* leaf) { return grad_of(tr, "blocks." + std::to_string(l) + "." + leaf); };
  if (layer_hi == g->L) {
    MGV_CHECK_CUDA(cudaMemsetAsync(tr->grads, 0, static_cast<size_t>(tr->total) * 4, s));
    // pad columns of the transposed operands must be zero (they are contracted over)
    if (Rp != R) {
      MGV_CHECK_CUDA(cudaMemsetAsync(tr->gT, 0, static_cast<size_t>(4) * C * Rp * 2, s));
      MGV_CHECK_CUDA(cudaMemsetAsync(tr->actT, 0, static_cast<size_t>(4) * C * Rp * 2, s));
    }
    // head: dW = dlogits^T f ; df = dlogits W_head ; ln_f backward -> dx
    MGV_TRY(train_grad_prep(GRAD_PREP_COPY, tr->dlogits, nullptr, R, V, Rp, nullptr, tr->gT, nullptr, none, 0, s));
    MGV_TRY(train_transpose(tr->f, R, C, C, Rp, tr->actT, s));
    MGV_TRY(gemm(tr->gT, tr->actT, V, C, Rp, EPI_F32, nullptr, grad_of(tr, "head.weight"), nullptr, s));
    MGV_TRY(gemm(tr->dlogits, tr->wheadT, R, C, V, EPI_F32, nullptr, tr->tmp32, nullptr, s));
    MGV_TRY(train_layernorm_bwd(tr->tmp32, tr->xf, g->lnf_w, R, C, tr->dx, false, grad_of(tr, "ln_f.weight"), grad_of(tr, "ln_f.bias"), s));
    g->launches += 5;
  }
  for (int l = layer_hi - 1; l >= layer_lo; --l) {
    const GptLayer& w = g->layers[l];
    TrainLayerActs& a = tr->acts[l];
    // ---- mlp branch: x2 = x1 + drop(h W2^T + b2)
    MGV_TRY(train_grad_prep(GRAD_PREP_DROP, tr->dx, nullptr, R, C, Rp, tr->g, tr->gT, gname(l, "mlp.2.bias"), tr->d_resid,
                            drop_stream_resid_mlp(l), s));
    MGV_TRY(train_transpose(a.h, R, 4 * C, 4 * C, Rp, tr->actT, s));
    MGV_TRY(gemm(tr->gT, tr->actT, C, 4 * C, Rp, EPI_F32, nullptr, gname(l, "mlp.2.weight"), nullptr, s));       // dW2 [C, 4C]
    MGV_TRY(gemm(tr->g, tr->wfc2T[l], R, 4 * C, C, EPI_BF16, nullptr, tr->wide, nullptr, s));                    // dh [R, 4C]
    MGV_TRY(train_grad_prep(GRAD_PREP_GELU, tr->wide, a.hpre, R, 4 * C, Rp, tr->g, tr->gT, gname(l, "mlp.0.bias"), none, 0, s));
    MGV_TRY(train_transpose(a.c, R, C, C, Rp, tr->actT, s));
    MGV_TRY(gemm(tr->gT, tr->actT, 4 * C, C, Rp, EPI_F32, nullptr, gname(l, "mlp.0.weight"), nullptr, s));       // dW1 [4C, C]
    MGV_TRY(gemm(tr->g, tr->wfc1T[l], R, C, 4 * C, EPI_F32, nullptr, tr->tmp32, nullptr, s));                    // dc [R, C]
    MGV_TRY(train_layernorm_bwd(tr->tmp32, a.x1, w.ln2_w, R, C, tr->dx, true, gname(l, "ln2.weight"), gname(l, "ln2.bias"), s));
    // ---- attention branch: x1 = x0 + drop(y Wproj^T + b)
    MGV_TRY(train_grad_prep(GRAD_PREP_DROP, tr->dx, nullptr, R, C, Rp, tr->g, tr->gT, gname(l, "attn.proj.bias"), tr->d_resid,
                            drop_stream_resid_attn(l), s));
    MGV_TRY(train_transpose(a.y, R, C, C, Rp, tr->actT, s));
    MGV_TRY(gemm(tr->gT, tr->actT, C, C, Rp, EPI_F32, nullptr, gname(l, "attn.proj.weight"), nullptr, s));       // dWproj
    MGV_TRY(gemm(tr->g, tr->wprojT[l], R, C, C, EPI_BF16, nullptr, tr->wide, nullptr, s));                       // dy [R, C] bf16
    MGV_TRY(train_attn_bwd(a.qkv, a.y, tr->wide, a.lse2, B, T, g->nh, g->cfg.n_unmasked, tr->dqkv, tr->delta, tr->d_attn,
                           drop_stream_attn(l), s));
    // q|k|v biases are adjacent in the flat layout (query, key, value): one [3C] column sum
    MGV_TRY(train_grad_prep(GRAD_PREP_COPY, tr->dqkv, nullptr, R, 3 * C, Rp, nullptr, tr->gT, gname(l, "attn.query.bias"), none, 0, s));
    MGV_TRY(train_transpose(a.a, R, C, C, Rp, tr->actT, s));
    MGV_TRY(gemm(tr->gT, tr->actT, 3 * C, C, Rp, EPI_F32, nullptr, gname(l, "attn.query.weight"), nullptr, s));  // dWqkv [3C, C]
    MGV_TRY(gemm(tr->dqkv, tr->wqkvT[l], R, C, 3 * C, EPI_F32, nullptr, tr->tmp32, nullptr, s));                 // da [R, C]
    MGV_TRY(train_layernorm_bwd(tr->tmp32, a.x0, w.ln1_w, R, C, tr->dx, true, gname(l, "ln1.weight"), gname(l, "ln1.bias"), s));
    g->launches += 20;
  }
  if (layer_lo == 0) {
    MGV_TRY(train_embed_bwd(tr->dx, tr->idx, B, T, tr->t, tr->cls, tr->m, C, grad_of(tr, "tok_emb.weight"), grad_of(tr, "pos_emb"),
                            g->cfg.class_size > 0 ? grad_of(tr, "embedder.weight") : nullptr, tr->d_embd, s));
    g->launches += 1;
  }
  return MGV_OK;
}

// torch.optim.AdamW step over the flat buffers (m, v: caller-owned fp32 state of the same length); refreshes the handle's
// inference copies and the transposed weights.  grad_scale multiplies the gradients first (1 / world_size after a summing
// all-reduce).
int gpt_train_adamw(Gpt* g, float* m, float* v, float lr, float beta1, float beta2, float eps, float weight_decay,
                    long long step, float grad_scale, cudaStream_t s) {
  MGV_REQUIRE(g && g->train && g->train->params && g->train->d_segs, "gpt_train_adamw: call mgv_gpt_train_bind first");
  GptTrain* tr = g->train;
  MGV_TRY(train_adamw(tr->params, tr->grads, m, v, tr->d_segs, tr->d_chunks, tr->n_chunks, lr, beta1, beta2, eps, weight_decay,
                      step, grad_scale, s));
  g->fold_dirty = true;       // decode-chain fold vectors depend on the weights
  return refresh_transposed(g, tr, s);
}

}  // namespace mgv
