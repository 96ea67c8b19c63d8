// VQVAE encoder / decoder handle: packed bf16 weights, NHWC bf16 activations, every conv
// and 1x1 projection through the tcgen05 implicit-GEMM kernel (gemm_tc.cu), GroupNorm
// statistics fused into the producing conv's epilogue.
// reference: vqvae/big_model_attn_gan.py  Encoder :190-282, Decoder :291-392,
// ResnetBlock :75-135, AttnBlock :397-450, LitVQVAE.encode/decode :604-614,
// module constants :521-530.
#include <map>
#include <string>
#include <vector>
#include <stdlib.h>
#include "gemm_tc.cuh"
#include "vqvae.cuh"
#include "vqvae_kernels.cuh"

namespace mgv {

namespace {

constexpr int CH = 128;
constexpr int NUM_RES = 5;
const int CH_MULT[NUM_RES] = {1, 1, 2, 2, 4};
constexpr int NUM_RES_BLOCKS = 2;
constexpr int Z_CH = 256;
constexpr int MEL_H = 80, MEL_W = 848;
constexpr int LAT_H = 5, LAT_W = 53;
constexpr int MAX_TILES_PER_IMAGE = 80 * 7;   // 80x848 output: 7 x-tiles of 128 pixels per row

struct Conv {   // 3x3 or 1x1 convolution, weights [Cout][taps][Cin] bf16
  __nv_bfloat16* w = nullptr;
  __nv_bfloat16* w_up = nullptr;   // Upsample convs only: phase weights [4][Cout][2x2 taps][Cin] (vqvae_upsample_phase_weights)
  float* b = nullptr;
  int cin = 0, cout = 0, k = 0;
};
struct GN {
  float *w = nullptr, *b = nullptr;
  int c = 0;
};
struct ResBlock {
  GN n1, n2;
  Conv c1, c2, nin;
  bool has_nin = false;
};
struct Attn {
  GN n;
  __nv_bfloat16* wqkv = nullptr;  // [3C][C]
  float* bqkv = nullptr;          // [3C]
  Conv proj;
};

__global__ void cvt_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}
// (1, C, 3, 3) -> [9][C] fp32
__global__ void repack_convout_kernel(const float* __restrict__ in, int C, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 9 * C) {
    const int c = i % C, tap = i / C;
    out[i] = in[c * 9 + tap];
  }
}

}  // namespace

// how a state_dict tensor is stored in the handle
enum SlotKind { SLOT_F32, SLOT_CONV3, SLOT_CONV3_UP, SLOT_CONV1, SLOT_QKV_W, SLOT_QKV_B, SLOT_CONVOUT_W, SLOT_IGNORE };
struct Slot {
  SlotKind kind;
  void* dst;
  long long numel;
  int a, b;       // conv: cout, cin ; qkv: part, C
  bool loaded;
  void* dst2 = nullptr;   // SLOT_CONV3_UP: the phase weights
};

struct Vqvae {
  int K, D;
  // encoder
  float *enc_conv_in_w = nullptr, *enc_conv_in_b = nullptr;   // fp32 (128,1,3,3)
  ResBlock enc_down[NUM_RES][NUM_RES_BLOCKS];
  Attn enc_down_attn[NUM_RES_BLOCKS];                          // level 4 only
  Conv enc_downsample[NUM_RES - 1];
  ResBlock enc_mid1, enc_mid2;
  Attn enc_mid_attn;
  GN enc_norm_out;
  Conv enc_conv_out, quant_conv;
  // codebook + decoder
  float* codebook = nullptr;
  float *pq_w = nullptr, *pq_b = nullptr;   // post_quant_conv fp32 (for the gather table)
  Conv post_quant;                           // bf16 copy (decode(quant) path)
  __nv_bfloat16* table = nullptr;            // [K][256]
  bool table_valid = false;
  Conv dec_conv_in;
  ResBlock dec_mid1, dec_mid2;
  Attn dec_mid_attn;
  ResBlock dec_up[NUM_RES][NUM_RES_BLOCKS + 1];
  Attn dec_up_attn[NUM_RES_BLOCKS + 1];     // level 4 only
  Conv dec_upsample[NUM_RES];               // levels 1..4
  GN dec_norm_out;
  float *dec_conv_out_w = nullptr, *dec_conv_out_b = nullptr;  // [9][128] fp32, [1]

  std::map<std::string, Slot> slots;
  std::vector<void*> allocs;
  // workspaces
  int ws_B = 0;
  __nv_bfloat16* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* f32buf = nullptr;     // [B,265,256] fp32 (encoder output before NCHW transpose)
  float* gn_slab = nullptr;    // per-GroupNorm finalized statistics slots [slot][B][32][{mean, rstd}]
  float* gn_part = nullptr;    // per-tile partial sums of the tensor produced last [B][tiles][32][2] (reused)
  int gn_slots = 0;
  int* flag = nullptr;
  long long launches = 0;
  // CUDA graph of the decoder (vqvae_decode) and the staging buffers its captured pointers refer to
  cudaGraph_t dec_graph = nullptr;
  cudaGraphExec_t dec_exec = nullptr;
  int dec_graph_B = 0, eager_B = 0;
  long long dec_graph_launches = 0;
  cudaStream_t cap_stream = nullptr;
  long long* idx_stage = nullptr;   // [ws_B, 265]
  float* mel_stage = nullptr;       // [ws_B, 80, 848]
};

namespace {

template <typename T>
T* dalloc(Vqvae* v, size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
  v->allocs.push_back(p);
  return static_cast<T*>(p);
}

void reg_f32(Vqvae* v, const std::string& name, float** dst, long long n) {
  *dst = dalloc<float>(v, n);
  v->slots[name] = Slot{SLOT_F32, *dst, n, 0, 0, false};
}
void reg_conv(Vqvae* v, const std::string& p, Conv* c, int cin, int cout, int k) {
  c->cin = cin; c->cout = cout; c->k = k;
  c->w = dalloc<__nv_bfloat16>(v, static_cast<size_t>(cout) * cin * k * k);
  c->b = dalloc<float>(v, cout);
  v->slots[p + ".weight"] = Slot{k == 3 ? SLOT_CONV3 : SLOT_CONV1, c->w, static_cast<long long>(cout) * cin * k * k, cout, cin, false};
  v->slots[p + ".bias"] = Slot{SLOT_F32, c->b, cout, 0, 0, false};
}
void reg_gn(Vqvae* v, const std::string& p, GN* g, int c) {
  g->c = c;
  reg_f32(v, p + ".weight", &g->w, c);
  reg_f32(v, p + ".bias", &g->b, c);
}
void reg_res(Vqvae* v, const std::string& p, ResBlock* r, int cin, int cout) {
  reg_gn(v, p + ".norm1", &r->n1, cin);
  reg_conv(v, p + ".conv1", &r->c1, cin, cout, 3);
  reg_gn(v, p + ".norm2", &r->n2, cout);
  reg_conv(v, p + ".conv2", &r->c2, cout, cout, 3);
  r->has_nin = cin != cout;
  if (r->has_nin) reg_conv(v, p + ".nin_shortcut", &r->nin, cin, cout, 1);
}
void reg_attn(Vqvae* v, const std::string& p, Attn* a, int c) {
  reg_gn(v, p + ".norm", &a->n, c);
  a->wqkv = dalloc<__nv_bfloat16>(v, static_cast<size_t>(3) * c * c);
  a->bqkv = dalloc<float>(v, 3 * c);
  const char* names[3] = {".q", ".k", ".v"};
  for (int i = 0; i < 3; ++i) {
    v->slots[p + names[i] + ".weight"] = Slot{SLOT_QKV_W, a->wqkv + static_cast<size_t>(i) * c * c, static_cast<long long>(c) * c, i, c, false};
    v->slots[p + names[i] + ".bias"] = Slot{SLOT_QKV_B, a->bqkv + i * c, c, i, c, false};
  }
  reg_conv(v, p + ".proj_out", &a->proj, c, c, 1);
}

std::string fmt(const char* f, int a, int b = 0) {
  char buf[128];
  snprintf(buf, sizeof(buf), f, a, b);
  return buf;
}

void register_all(Vqvae* v) {
  // ---- encoder (Encoder.__init__ :190-251)
  v->enc_conv_in_w = dalloc<float>(v, CH * 9);
  v->slots["_encoder.conv_in.weight"] = Slot{SLOT_F32, v->enc_conv_in_w, CH * 9, 0, 0, false};
  reg_f32(v, "_encoder.conv_in.bias", &v->enc_conv_in_b, CH);
  int block_in = CH;
  for (int lvl = 0; lvl < NUM_RES; ++lvl) {
    const int in_mult = lvl == 0 ? 1 : CH_MULT[lvl - 1];
    block_in = CH * in_mult;
    const int block_out = CH * CH_MULT[lvl];
    for (int b = 0; b < NUM_RES_BLOCKS; ++b) {
      reg_res(v, fmt("_encoder.down.%d.block.%d", lvl, b), &v->enc_down[lvl][b], block_in, block_out);
      block_in = block_out;
      if (lvl == NUM_RES - 1) reg_attn(v, fmt("_encoder.down.%d.attn.%d", lvl, b), &v->enc_down_attn[b], block_in);
    }
    if (lvl != NUM_RES - 1) reg_conv(v, fmt("_encoder.down.%d.downsample.conv", lvl), &v->enc_downsample[lvl], block_in, block_in, 3);
  }
  reg_res(v, "_encoder.mid.block_1", &v->enc_mid1, block_in, block_in);
  reg_attn(v, "_encoder.mid.attn_1", &v->enc_mid_attn, block_in);
  reg_res(v, "_encoder.mid.block_2", &v->enc_mid2, block_in, block_in);
  reg_gn(v, "_encoder.norm_out", &v->enc_norm_out, block_in);
  reg_conv(v, "_encoder.conv_out", &v->enc_conv_out, block_in, Z_CH, 3);
  reg_conv(v, "quant_conv", &v->quant_conv, Z_CH, v->D, 1);
  // ---- codebook + post_quant_conv (kept in fp32 too, for the fused gather table)
  reg_f32(v, "_vq_vae._embedding.weight", &v->codebook, static_cast<long long>(v->K) * v->D);
  v->pq_w = dalloc<float>(v, static_cast<size_t>(Z_CH) * v->D);
  v->pq_b = dalloc<float>(v, Z_CH);
  reg_conv(v, "post_quant_conv", &v->post_quant, v->D, Z_CH, 1);
  v->table = dalloc<__nv_bfloat16>(v, static_cast<size_t>(v->K) * Z_CH);
  // ---- decoder (Decoder.__init__ :291-359)
  block_in = CH * CH_MULT[NUM_RES - 1];
  reg_conv(v, "_decoder.conv_in", &v->dec_conv_in, Z_CH, block_in, 3);
  reg_res(v, "_decoder.mid.block_1", &v->dec_mid1, block_in, block_in);
  reg_attn(v, "_decoder.mid.attn_1", &v->dec_mid_attn, block_in);
  reg_res(v, "_decoder.mid.block_2", &v->dec_mid2, block_in, block_in);
  for (int lvl = NUM_RES - 1; lvl >= 0; --lvl) {
    const int block_out = CH * CH_MULT[lvl];
    for (int b = 0; b < NUM_RES_BLOCKS + 1; ++b) {
      reg_res(v, fmt("_decoder.up.%d.block.%d", lvl, b), &v->dec_up[lvl][b], block_in, block_out);
      block_in = block_out;
      if (lvl == NUM_RES - 1) reg_attn(v, fmt("_decoder.up.%d.attn.%d", lvl, b), &v->dec_up_attn[b], block_in);
    }
    if (lvl != 0) {
      Conv* uc = &v->dec_upsample[lvl];
      reg_conv(v, fmt("_decoder.up.%d.upsample.conv", lvl), uc, block_in, block_in, 3);
      uc->w_up = dalloc<__nv_bfloat16>(v, static_cast<size_t>(16) * block_in * block_in);
      Slot& sl = v->slots[fmt("_decoder.up.%d.upsample.conv.weight", lvl)];
      sl.kind = SLOT_CONV3_UP;
      sl.dst2 = uc->w_up;
    }
  }
  reg_gn(v, "_decoder.norm_out", &v->dec_norm_out, block_in);
  v->dec_conv_out_w = dalloc<float>(v, 9 * block_in);
  v->slots["_decoder.conv_out.weight"] = Slot{SLOT_CONVOUT_W, v->dec_conv_out_w, 9LL * block_in, 0, block_in, false};
  reg_f32(v, "_decoder.conv_out.bias", &v->dec_conv_out_b, 1);
}

int ensure_ws(Vqvae* v, int B) {
  if (B <= v->ws_B) return MGV_OK;
  for (auto& p : v->buf) { cudaFree(p); p = nullptr; }
  cudaFree(v->f32buf); v->f32buf = nullptr;
  cudaFree(v->gn_slab); v->gn_slab = nullptr;
  cudaFree(v->gn_part); v->gn_part = nullptr;
  cudaFree(v->idx_stage); v->idx_stage = nullptr;
  cudaFree(v->mel_stage); v->mel_stage = nullptr;
  if (v->dec_exec) {   // the captured launches point into the buffers that were just freed
    cudaGraphExecDestroy(v->dec_exec);
    cudaGraphDestroy(v->dec_graph);
    v->dec_exec = nullptr;
    v->dec_graph = nullptr;
  }
  v->eager_B = 0;
  v->ws_B = 0;
  const size_t act = static_cast<size_t>(B) * MEL_H * MEL_W * CH;  // largest activation (elements)
  for (auto& p : v->buf) MGV_CHECK_CUDA(cudaMalloc(&p, act * 2));
  MGV_CHECK_CUDA(cudaMalloc(&v->f32buf, static_cast<size_t>(B) * LAT_H * LAT_W * (v->D > Z_CH ? v->D : Z_CH) * 4));
  v->gn_slots = 96;
  MGV_CHECK_CUDA(cudaMalloc(&v->gn_slab, static_cast<size_t>(v->gn_slots) * B * 64 * 4));
  MGV_CHECK_CUDA(cudaMalloc(&v->gn_part, static_cast<size_t>(B) * MAX_TILES_PER_IMAGE * 64 * 4));
  MGV_CHECK_CUDA(cudaMalloc(&v->idx_stage, static_cast<size_t>(B) * LAT_H * LAT_W * 8));
  MGV_CHECK_CUDA(cudaMalloc(&v->mel_stage, static_cast<size_t>(B) * MEL_H * MEL_W * 4));
  v->ws_B = B;
  return MGV_OK;
}

int check_loaded(const Vqvae* v, const char* prefix_a, const char* prefix_b, const char* prefix_c) {
  for (const auto& kv : v->slots) {
    const std::string& n = kv.first;
    const bool wanted = n.rfind(prefix_a, 0) == 0 || n.rfind(prefix_b, 0) == 0 || (prefix_c && n.rfind(prefix_c, 0) == 0);
    if (wanted && !kv.second.loaded) {
      set_error("vqvae: weight '%s' was never loaded", n.c_str());
      return MGV_ERR_STATE;
    }
  }
  return MGV_OK;
}

// ---- execution context: activation buffers + GroupNorm statistics slots
struct Ctx {
  Vqvae* v;
  int B;
  cudaStream_t s;
  int next_slot = 0;
  float* new_stats() {
    float* p = v->gn_slab + static_cast<size_t>(next_slot) * B * 64;
    next_slot++;
    return p;
  }
};

int stats_of(struct Ctx& c, const __nv_bfloat16* x, int HW, int C, float* stats_out);

int conv_stages() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MGV_CONV_STAGES");
    v = e ? atoi(e) : 3;
    if (v < 1) v = 1;
  }
  return v;
}

// 3x3 conv (stride 1 pad 1, or the Downsample variant) with optional residual and fused GN statistics of the output
int conv3(Ctx& c, const Conv& w, const __nv_bfloat16* x, int Hin, int Win, int stride, __nv_bfloat16* out,
          const __nv_bfloat16* resid, float* stats_out) {
  GemmArgs a;
  a.a_mode = A_CONV3x3;
  a.A = x; a.B = w.w;
  a.n_img = c.B; a.Hin = Hin; a.Win = Win; a.Cin = w.cin; a.stride = stride;
  if (stride == 1) { a.H = Hin; a.W = Win; a.pad = 1; }
  else { a.H = (Hin + 1 - 3) / 2 + 1; a.W = (Win + 1 - 3) / 2 + 1; a.pad = 0; }   // F.pad (0,1,0,1) + stride 2 (:151-159)
  a.M = c.B * a.H * a.W; a.N = w.cout; a.K = 9 * w.cin;
  a.epi = resid ? EPI_BF16_RESID : EPI_BF16;
  a.bias = w.b; a.out = out; a.resid = resid;
  a.bn = (w.cout % 256 == 0 && !getenv("MGV_NO_BN256")) ? 256 : 128;   // 128x256 tiles move 25% fewer L2 bytes per FLOP
  a.max_stages = conv_stages();   // 3 stages = 96 KB: two CTAs per SM, one's epilogue overlaps the other's MMAs
  static const bool unfused = getenv("MGV_UNFUSED_GNSTATS") != nullptr;   // debugging aid
  if (stats_out && !unfused) { a.gn_sum = c.v->gn_part; a.gn_group_ch = w.cout / 32; }
  a.stream = c.s;
  c.v->launches++;
  MGV_TRY(gemm_bf16_tc(a));
  if (stats_out && unfused) return stats_of(c, out, a.H * a.W, w.cout, stats_out);
  if (stats_out) {
    // tiles per image of the conv kernel's grid (see fill_params in gemm_tc.cu)
    const int wb = conv_tile_width(a.H, a.W);
    const int tiles = ceil_div(a.W, wb) * ceil_div(a.H, 128 / wb);
    MGV_REQUIRE(tiles <= MAX_TILES_PER_IMAGE, "conv3: %d tiles per image exceed the statistics scratch", tiles);
    MGV_TRY(vqvae_gn_finalize(c.v->gn_part, c.B, tiles, static_cast<float>(a.H) * a.W * (w.cout / 32), stats_out, c.s));
    c.v->launches++;
  }
  return MGV_OK;
}

// Upsample.forward (:182-186): nearest 2x + 3x3 conv, computed as four 2x2 convolutions over the LOW-res input (one per
// output phase (py, px), weights pre-summed at load time): 16 instead of 36 multiply-adds per input pixel and channel
// pair, and the 4x larger upsampled tensor is never written or read.  x: B x H x W x cin; out: B x 2H x 2W x cout.
int conv_up(Ctx& c, const Conv& w, const __nv_bfloat16* x, int H, int W, __nv_bfloat16* out, float* stats_out) {
  const int wb = conv_tile_width(H, W);
  const int tiles = ceil_div(W, wb) * ceil_div(H, 128 / wb);   // per image and phase (fill_params in gemm_tc.cu)
  MGV_REQUIRE(4 * tiles <= MAX_TILES_PER_IMAGE, "conv_up: %d tiles per image exceed the statistics scratch", 4 * tiles);
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    GemmArgs a;
    a.a_mode = A_CONV3x3;
    a.A = x; a.B = w.w_up + static_cast<size_t>(ph) * w.cout * 4 * w.cin;
    a.n_img = c.B; a.Hin = H; a.Win = W; a.Cin = w.cin; a.stride = 1;
    a.H = H; a.W = W;
    a.taps_x = 2; a.pad = 1 - px; a.pad_y = 1 - py;
    a.out_scale = 2; a.out_oy = py; a.out_ox = px;
    a.M = c.B * H * W; a.N = w.cout; a.K = 4 * w.cin;
    a.epi = EPI_BF16;
    a.bias = w.b; a.out = out;
    a.bn = (w.cout % 256 == 0 && !getenv("MGV_NO_BN256")) ? 256 : 128;
    a.max_stages = conv_stages();
    if (stats_out) { a.gn_sum = c.v->gn_part; a.gn_group_ch = w.cout / 32; a.gn_tiles_img = 4 * tiles; a.gn_tile_base = ph * tiles; }
    a.stream = c.s;
    c.v->launches++;
    MGV_TRY(gemm_bf16_tc(a));
  }
  if (stats_out) {
    MGV_TRY(vqvae_gn_finalize(c.v->gn_part, c.B, 4 * tiles, 4.0f * H * W * (w.cout / 32), stats_out, c.s));
    c.v->launches++;
  }
  return MGV_OK;
}

// 1x1 conv == GEMM over NHWC pixels
int conv1(Ctx& c, const __nv_bfloat16* wmat, const float* bias, int cin, int cout, const __nv_bfloat16* x, long long rows,
          void* out, int epi, const void* resid) {
  GemmArgs a;
  a.A = x; a.B = wmat; a.M = static_cast<int>(rows); a.N = cout; a.K = cin;
  a.epi = epi; a.bias = bias; a.out = out; a.resid = resid;
  a.bn = (cout % 256 == 0 && !getenv("MGV_NO_BN256")) ? 256 : 128;
  a.max_stages = conv_stages();
  a.stream = c.s;
  c.v->launches++;
  return gemm_bf16_tc(a);
}

// GroupNorm apply (+ optional swish) from accumulated statistics
int gn(Ctx& c, const GN& g, const __nv_bfloat16* x, const float* stats, int HW, int do_swish, __nv_bfloat16* y) {
  c.v->launches += 1;
  return vqvae_gn_apply(x, stats, g.w, g.b, c.B, HW, g.c, do_swish, y, c.s);
}

// statistics of a tensor that was not produced by a conv epilogue (attention output, encoder conv_in)
int stats_of(Ctx& c, const __nv_bfloat16* x, int HW, int C, float* stats_out) {
  int tiles = 0;
  MGV_TRY(vqvae_gn_stats(x, c.B, HW, C, c.v->gn_part, &tiles, c.s));
  MGV_REQUIRE(tiles <= MAX_TILES_PER_IMAGE, "gn_stats: too many chunks");
  c.v->launches += 2;
  return vqvae_gn_finalize(c.v->gn_part, c.B, tiles, static_cast<float>(HW) * (C / 32), stats_out, c.s);
}

// ResnetBlock.forward (:114-135).  x (with statistics x_stats) -> out (statistics out_stats).
// tmp_a / tmp_b / tmp_s are scratch buffers distinct from x and out.
int resblock(Ctx& c, const ResBlock& r, const __nv_bfloat16* x, const float* x_stats, int H, int W, __nv_bfloat16* out,
             float* out_stats, __nv_bfloat16* tmp_a, __nv_bfloat16* tmp_b, __nv_bfloat16* tmp_s) {
  const int HW = H * W;
  MGV_TRY(gn(c, r.n1, x, x_stats, HW, 1, tmp_a));
  float* st1 = c.new_stats();
  MGV_TRY(conv3(c, r.c1, tmp_a, H, W, 1, tmp_b, nullptr, st1));
  MGV_TRY(gn(c, r.n2, tmp_b, st1, HW, 1, tmp_a));
  const __nv_bfloat16* shortcut = x;
  if (r.has_nin) {
    MGV_TRY(conv1(c, r.nin.w, r.nin.b, r.nin.cin, r.nin.cout, x, static_cast<long long>(c.B) * HW, tmp_s, EPI_BF16, nullptr));
    shortcut = tmp_s;
  }
  MGV_TRY(conv3(c, r.c2, tmp_a, H, W, 1, out, shortcut, out_stats));
  return MGV_OK;
}

// AttnBlock.forward (:425-450).  x -> out (= x + proj(attn)); statistics of out computed by a separate pass.
int attnblock(Ctx& c, const Attn& a, const __nv_bfloat16* x, const float* x_stats, int H, int W, __nv_bfloat16* out,
              float* out_stats, __nv_bfloat16* tmp_a, __nv_bfloat16* tmp_b) {
  const int HW = H * W, C = a.n.c;
  const long long rows = static_cast<long long>(c.B) * HW;
  MGV_TRY(gn(c, a.n, x, x_stats, HW, 0, tmp_a));
  MGV_TRY(conv1(c, a.wqkv, a.bqkv, C, 3 * C, tmp_a, rows, tmp_b, EPI_BF16, nullptr));   // q | k | v
  MGV_TRY(vqvae_spatial_attention(tmp_b, c.B, HW, C, tmp_a, c.s));
  MGV_TRY(conv1(c, a.proj.w, a.proj.b, C, C, tmp_a, rows, out, EPI_BF16_RESID, x));
  if (out_stats) MGV_TRY(stats_of(c, out, HW, C, out_stats));
  c.v->launches += 1;
  return MGV_OK;
}

}  // namespace

int vqvae_create(int num_embeddings, int embedding_dim, Vqvae** out) {
  MGV_REQUIRE(out, "vqvae_create: null");
  MGV_TRY(check_device());
  MGV_REQUIRE(embedding_dim % 64 == 0 && embedding_dim >= 64 && embedding_dim <= 1024, "vqvae: embedding_dim=%d unsupported",
              embedding_dim);
  MGV_REQUIRE(embedding_dim % 32 == 0, "vqvae: embedding_dim");
  MGV_REQUIRE(num_embeddings >= 1, "vqvae: num_embeddings=%d", num_embeddings);
  Vqvae* v = new Vqvae();
  v->K = num_embeddings;
  v->D = embedding_dim;
  register_all(v);
  cudaMalloc(&v->flag, sizeof(int));
  cudaMemset(v->flag, 0, sizeof(int));
  for (void* p : v->allocs)
    if (!p) {
      set_error("vqvae_create: cudaMalloc failed");
      vqvae_destroy(v);
      return MGV_ERR_CUDA;
    }
  if (cudaGetLastError() != cudaSuccess) {
    set_error("vqvae_create: CUDA error");
    vqvae_destroy(v);
    return MGV_ERR_CUDA;
  }
  *out = v;
  return MGV_OK;
}

int vqvae_destroy(Vqvae* v) {
  if (!v) return MGV_OK;
  for (void* p : v->allocs) cudaFree(p);
  for (auto& p : v->buf) cudaFree(p);
  cudaFree(v->f32buf);
  cudaFree(v->gn_slab);
  cudaFree(v->gn_part);
  cudaFree(v->flag);
  cudaFree(v->idx_stage);
  cudaFree(v->mel_stage);
  if (v->dec_exec) cudaGraphExecDestroy(v->dec_exec);
  if (v->dec_graph) cudaGraphDestroy(v->dec_graph);
  if (v->cap_stream) cudaStreamDestroy(v->cap_stream);
  delete v;
  return MGV_OK;
}

int vqvae_load_weight(Vqvae* v, const char* name, const float* src, long long numel, cudaStream_t s) {
  MGV_REQUIRE(v && name && src, "vqvae_load_weight: null");
  const std::string k(name);
  if (k.rfind("discriminator.", 0) == 0) return MGV_OK;  // GAN critic: never on the inference path
  auto it = v->slots.find(k);
  if (it == v->slots.end()) {
    set_error("vqvae_load_weight: unknown tensor name '%s'", name);
    return MGV_ERR_INVALID;
  }
  Slot& sl = it->second;
  MGV_REQUIRE(numel == sl.numel, "vqvae_load_weight(%s): numel %lld != expected %lld", name, numel, sl.numel);
  const int blocks = static_cast<int>(numel > 256 * 2048 ? 2048 : (numel + 255) / 256);
  switch (sl.kind) {
    case SLOT_F32:
      MGV_CHECK_CUDA(cudaMemcpyAsync(sl.dst, src, numel * 4, cudaMemcpyDeviceToDevice, s));
      break;
    case SLOT_CONV3:
    case SLOT_CONV3_UP:
      MGV_TRY(vqvae_repack_conv_weight(src, sl.a, sl.b, 3, 3, static_cast<__nv_bfloat16*>(sl.dst), s));
      if (sl.kind == SLOT_CONV3_UP)
        MGV_TRY(vqvae_upsample_phase_weights(src, sl.a, sl.b, static_cast<__nv_bfloat16*>(sl.dst2), s));
      break;
    case SLOT_CONV1:
    case SLOT_QKV_W:
      cvt_bf16_kernel<<<blocks, 256, 0, s>>>(src, static_cast<__nv_bfloat16*>(sl.dst), numel);
      MGV_CHECK_CUDA(cudaGetLastError());
      break;
    case SLOT_QKV_B:
      MGV_CHECK_CUDA(cudaMemcpyAsync(sl.dst, src, numel * 4, cudaMemcpyDeviceToDevice, s));
      break;
    case SLOT_CONVOUT_W:
      repack_convout_kernel<<<ceil_div(9 * sl.b, 256), 256, 0, s>>>(src, sl.b, static_cast<float*>(sl.dst));
      MGV_CHECK_CUDA(cudaGetLastError());
      break;
    default: break;
  }
  sl.loaded = true;
  if (k == "post_quant_conv.weight") {
    MGV_CHECK_CUDA(cudaMemcpyAsync(v->pq_w, src, numel * 4, cudaMemcpyDeviceToDevice, s));
    v->table_valid = false;
  } else if (k == "post_quant_conv.bias") {
    MGV_CHECK_CUDA(cudaMemcpyAsync(v->pq_b, src, numel * 4, cudaMemcpyDeviceToDevice, s));
    v->table_valid = false;
  } else if (k == "_vq_vae._embedding.weight") {
    v->table_valid = false;
  }
  return MGV_OK;
}

// every launch of the decoder for B images, enqueued on s (the gather table is valid; no host synchronisation inside, so
// the sequence can be captured into a CUDA graph)
static int decode_body(Vqvae* v, const long long* idx, const float* quant_bchw, int B, float* mel_out, cudaStream_t s) {
  Ctx c{v, B, s};
  __nv_bfloat16 *h = v->buf[0], *o = v->buf[1], *ta = v->buf[2], *tb = v->buf[3], *ts = v->buf[4];
  const long long lat_rows = static_cast<long long>(B) * LAT_H * LAT_W;
  // ---- z_q -> post_quant_conv (fused into a table lookup when decoding codes)
  if (idx) {
    MGV_TRY(vqvae_gather_rows(idx, v->table, lat_rows, Z_CH, v->K, ta, v->flag, s));
    v->launches++;
  } else {
    MGV_REQUIRE(v->D == Z_CH || v->D % 64 == 0, "vqvae_decode: embedding_dim");
    MGV_TRY(vqvae_nchw_f32_to_nhwc_bf16(quant_bchw, B, v->D, LAT_H * LAT_W, tb, s));
    MGV_TRY(conv1(c, v->post_quant.w, v->post_quant.b, v->D, Z_CH, tb, lat_rows, ta, EPI_BF16, nullptr));
    v->launches++;
  }
  int H = LAT_H, W = LAT_W;
  float* st = c.new_stats();
  MGV_TRY(conv3(c, v->dec_conv_in, ta, H, W, 1, h, nullptr, st));                        // :369
  // ---- middle (:372-374)
  float* st2 = c.new_stats();
  MGV_TRY(resblock(c, v->dec_mid1, h, st, H, W, o, st2, ta, tb, ts));
  std::swap(h, o); st = st2;
  st2 = c.new_stats();
  MGV_TRY(attnblock(c, v->dec_mid_attn, h, st, H, W, o, st2, ta, tb));
  std::swap(h, o); st = st2;
  st2 = c.new_stats();
  MGV_TRY(resblock(c, v->dec_mid2, h, st, H, W, o, st2, ta, tb, ts));
  std::swap(h, o); st = st2;
  // ---- upsampling (:377-383)
  for (int lvl = NUM_RES - 1; lvl >= 0; --lvl) {
    for (int b = 0; b < NUM_RES_BLOCKS + 1; ++b) {
      st2 = c.new_stats();
      MGV_TRY(resblock(c, v->dec_up[lvl][b], h, st, H, W, o, st2, ta, tb, ts));
      std::swap(h, o); st = st2;
      if (lvl == NUM_RES - 1) {
        st2 = c.new_stats();
        MGV_TRY(attnblock(c, v->dec_up_attn[b], h, st, H, W, o, st2, ta, tb));
        std::swap(h, o); st = st2;
      }
    }
    if (lvl != 0) {
      const Conv& uc = v->dec_upsample[lvl];
      st2 = c.new_stats();
      static const bool phase_form = getenv("MGV_UPSAMPLE_EXPLICIT") == nullptr;
      if (phase_form) {
        MGV_TRY(conv_up(c, uc, h, H, W, o, st2));
        H *= 2; W *= 2;
      } else {   // the literal form: materialise the upsampled tensor, then the 3x3 conv (kept for A/B measurements)
        MGV_TRY(vqvae_upsample2x(h, B, H, W, uc.cin, ta, s));
        H *= 2; W *= 2;
        MGV_TRY(conv3(c, uc, ta, H, W, 1, o, nullptr, st2));
        v->launches++;
      }
      std::swap(h, o); st = st2;
    }
  }
  MGV_REQUIRE(c.next_slot <= v->gn_slots, "vqvae_decode: statistics slots exhausted");
  // ---- end (:389-391)
  MGV_TRY(vqvae_norm_swish_conv_out(h, st, v->dec_norm_out.w, v->dec_norm_out.b, v->dec_conv_out_w, v->dec_conv_out_b, B,
                                    H, W, CH, mel_out, s));
  v->launches += 1;
  return MGV_OK;
}

// decode_to_img after code_reader (minGPT.py:515-528) / LitVQVAE.decode (:610-614)
//
// The decoder is ~150 launches, the first ~60 of which (5x53 and 10x106 levels) last 5-30 us each: launched one by one
// (two tensor-map encodes + a launch per GEMM) the host falls behind and the GPU idles ~1.6 ms of a 22 ms decode at
// B = 64.  The code path therefore replays a CUDA graph of the whole decoder: the first call at a batch size runs
// eagerly (function attributes, lazily built state), the second captures, later ones replay.  Indices and mels go
// through handle-owned staging buffers so that the captured pointers never change.
int vqvae_decode(Vqvae* v, const long long* idx, const float* quant_bchw, int B, float* mel_out, cudaStream_t s) {
  MGV_REQUIRE(v && mel_out && ((idx != nullptr) != (quant_bchw != nullptr)), "vqvae_decode: need exactly one of idx / quant");
  MGV_REQUIRE(B >= 0, "vqvae_decode: B=%d", B);
  MGV_TRY(check_loaded(v, "_decoder.", "post_quant_conv.", idx ? "_vq_vae." : nullptr));
  if (B == 0) return MGV_OK;
  v->launches = 0;
  MGV_TRY(ensure_ws(v, B));
  if (idx && !v->table_valid) {
    MGV_TRY(vqvae_build_gather_table(v->codebook, v->pq_w, v->pq_b, v->K, v->D, Z_CH, v->table, s));
    v->table_valid = true;
    v->launches++;
  }
  static const bool no_graph = getenv("MGV_VQVAE_NO_GRAPH") != nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  MGV_CHECK_CUDA(cudaStreamIsCapturing(s, &cap));
  const bool graph_ok = idx && !no_graph && cap == cudaStreamCaptureStatusNone;
  if (graph_ok && v->dec_exec == nullptr && v->eager_B == B) {
    // second call at this batch size: capture (on the handle's own stream -- the caller's may be the legacy stream)
    if (!v->cap_stream) MGV_CHECK_CUDA(cudaStreamCreateWithFlags(&v->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    MGV_CHECK_CUDA(cudaStreamBeginCapture(v->cap_stream, cudaStreamCaptureModeThreadLocal));
    const long long before = v->launches;
    const int rc = decode_body(v, v->idx_stage, nullptr, B, v->mel_stage, v->cap_stream);
    v->dec_graph_launches = v->launches - before;
    v->launches = before;
    const cudaError_t ce = cudaStreamEndCapture(v->cap_stream, &graph);
    if (rc != MGV_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    MGV_CHECK_CUDA(ce);
    MGV_CHECK_CUDA(cudaGraphInstantiate(&v->dec_exec, graph, 0));
    v->dec_graph = graph;
    v->dec_graph_B = B;
  }
  if (graph_ok && v->dec_exec != nullptr && v->dec_graph_B == B) {
    MGV_CHECK_CUDA(cudaMemcpyAsync(v->idx_stage, idx, static_cast<size_t>(B) * LAT_H * LAT_W * 8, cudaMemcpyDeviceToDevice, s));
    MGV_CHECK_CUDA(cudaGraphLaunch(v->dec_exec, s));
    MGV_CHECK_CUDA(cudaMemcpyAsync(mel_out, v->mel_stage, static_cast<size_t>(B) * MEL_H * MEL_W * 4, cudaMemcpyDeviceToDevice, s));
    v->launches += v->dec_graph_launches;
  } else {
    MGV_TRY(decode_body(v, idx, quant_bchw, B, mel_out, s));
    if (graph_ok) {
      if (v->dec_exec) {   // another batch size: drop the old graph, this size captures on its next call
        cudaGraphExecDestroy(v->dec_exec);
        cudaGraphDestroy(v->dec_graph);
        v->dec_exec = nullptr;
        v->dec_graph = nullptr;
      }
      v->eager_B = B;
    }
  }
  if (idx) {
    int flag = 0;
    MGV_CHECK_CUDA(cudaMemcpyAsync(&flag, v->flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    MGV_CHECK_CUDA(cudaStreamSynchronize(s));
    if (flag) {
      MGV_CHECK_CUDA(cudaMemsetAsync(v->flag, 0, sizeof(int), s));
      set_error("vqvae_decode: code index out of range [0, %d)", v->K);
      return MGV_ERR_INVALID;
    }
  }
  return MGV_OK;
}

// LitVQVAE.encode (:604-608)
int vqvae_encode(Vqvae* v, const float* mel, int B, float* z_out, cudaStream_t s) {
  MGV_REQUIRE(v && mel && z_out, "vqvae_encode: null");
  MGV_REQUIRE(B >= 0, "vqvae_encode: B=%d", B);
  MGV_TRY(check_loaded(v, "_encoder.", "quant_conv.", nullptr));
  if (B == 0) return MGV_OK;
  v->launches = 0;
  MGV_TRY(ensure_ws(v, B));
  Ctx c{v, B, s};
  __nv_bfloat16 *h = v->buf[0], *o = v->buf[1], *ta = v->buf[2], *tb = v->buf[3], *ts = v->buf[4];
  int H = MEL_H, W = MEL_W;
  MGV_TRY(vqvae_conv_in_1ch(mel, v->enc_conv_in_w, v->enc_conv_in_b, B, H, W, CH, h, s));        // :261
  float* st = c.new_stats();
  MGV_TRY(stats_of(c, h, H * W, CH, st));
  v->launches += 1;
  float* st2;
  for (int lvl = 0; lvl < NUM_RES; ++lvl) {
    for (int b = 0; b < NUM_RES_BLOCKS; ++b) {
      st2 = c.new_stats();
      MGV_TRY(resblock(c, v->enc_down[lvl][b], h, st, H, W, o, st2, ta, tb, ts));
      std::swap(h, o); st = st2;
      if (lvl == NUM_RES - 1) {
        st2 = c.new_stats();
        MGV_TRY(attnblock(c, v->enc_down_attn[b], h, st, H, W, o, st2, ta, tb));
        std::swap(h, o); st = st2;
      }
    }
    if (lvl != NUM_RES - 1) {
      st2 = c.new_stats();
      MGV_TRY(conv3(c, v->enc_downsample[lvl], h, H, W, 2, o, nullptr, st2));               // :156-159
      H = (H + 1 - 3) / 2 + 1; W = (W + 1 - 3) / 2 + 1;
      std::swap(h, o); st = st2;
    }
  }
  st2 = c.new_stats();
  MGV_TRY(resblock(c, v->enc_mid1, h, st, H, W, o, st2, ta, tb, ts));
  std::swap(h, o); st = st2;
  st2 = c.new_stats();
  MGV_TRY(attnblock(c, v->enc_mid_attn, h, st, H, W, o, st2, ta, tb));
  std::swap(h, o); st = st2;
  st2 = c.new_stats();
  MGV_TRY(resblock(c, v->enc_mid2, h, st, H, W, o, st2, ta, tb, ts));
  std::swap(h, o); st = st2;
  MGV_REQUIRE(c.next_slot <= v->gn_slots, "vqvae_encode: statistics slots exhausted");
  MGV_REQUIRE(H == LAT_H && W == LAT_W, "vqvae_encode: unexpected latent size %dx%d", H, W);
  // ---- end (:278-280) + quant_conv (:606)
  MGV_TRY(gn(c, v->enc_norm_out, h, st, H * W, 1, ta));
  MGV_TRY(conv3(c, v->enc_conv_out, ta, H, W, 1, tb, nullptr, nullptr));
  const long long rows = static_cast<long long>(B) * H * W;
  MGV_TRY(conv1(c, v->quant_conv.w, v->quant_conv.b, Z_CH, v->D, tb, rows, v->f32buf, EPI_F32, nullptr));
  MGV_TRY(vqvae_nhwc_f32_to_nchw_f32(v->f32buf, B, v->D, H * W, z_out, s));
  v->launches += 2;
  return MGV_OK;
}

long long vqvae_last_launches(const Vqvae* v) { return v ? v->launches : 0; }

}  // namespace mgv
