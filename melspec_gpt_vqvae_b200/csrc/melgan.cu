// MelGAN generator (mel -> waveform), the step after the token path's end.
//   reference: vocoder/modules.py:23-36 (ResnetBlock), :38-80 (Generator);
//              caller callbacks/GPT_callbacks.py:93-105 (_log_rec_audio), GPT_VAE_callbacks.py:84-92.
//
// fp32 throughout (the reference runs it in fp32 on a handful of logged clips; audio is compared sample by sample).
// Activations stay in the reference's (B, C, T) layout.  Three kernels:
//   conv1d_kernel<KS>   stride-1 convolution with reflection padding and dilation; LeakyReLU(0.2) folded into the
//                       operand load; an optional SECOND input with its own 1x1 weights folds a ResnetBlock's shortcut
//                       and its last 1x1 convolution into one launch:  y = Ws x + W2 lrelu(h) + (bs + b2).
//   convt1d_kernel      ConvTranspose1d(kernel 2r, stride r, padding r/2 + r%2): every output sample has exactly two
//                       taps; a thread's outputs are 16 samples apart, so they share the tap phase and its weights.
//   conv_out_kernel     LeakyReLU -> ReflectionPad1d(3) -> Conv1d(ngf, 1, 7) -> tanh.
// A CTA computes 64 output channels x 128 time steps (256 threads, 4 channels x 8 steps each) from shared-memory tiles
// of 8 input channels at a time: FP32-FMA bound.
#include <string>
#include <utility>
#include <vector>
#include "mgv_common.cuh"
#include "melgan.cuh"

namespace mgv {

namespace {

constexpr int MG_CO = 64;      // output channels per CTA
constexpr int MG_T = 128;      // time steps per CTA
constexpr int MG_CI = 8;       // input channels per shared-memory stage
constexpr int MG_THREADS = 256;

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }
// ReflectionPad1d index (pad < T): -1 -> 1, T -> T-2
__device__ __forceinline__ int reflect(int i, int T) {
  if (i < 0) i = -i;
  if (i >= T) i = 2 * T - 2 - i;
  return i;
}

struct ConvArgs {
  const float* x; int cin; int lrelu_x;        // (B, cin, T)
  const float* x2; int cin2; int lrelu_x2;     // optional second input of a 1x1 pair (same T)
  const float* w;                              // packed [cin_total][KS][cout]
  const float* bias;                           // [cout] (sum of both biases for a pair)
  float* out; int cout; int T; int dil;
};

template <int KS>
__global__ void __launch_bounds__(MG_THREADS)
conv1d_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) float mg_smem[];
  const int halo = (KS - 1) / 2 * a.dil;
  const int xw = MG_T + 2 * halo;                      // staged window width
  float* xs = mg_smem;                                 // [MG_CI][xw]
  float* ws = mg_smem + MG_CI * xw;                    // [MG_CI][KS][MG_CO]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int t0 = blockIdx.x * MG_T, co0 = blockIdx.y * MG_CO, b = blockIdx.z;
  const int cin_total = a.cin + a.cin2;
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
  for (int ci0 = 0; ci0 < cin_total; ci0 += MG_CI) {
    __syncthreads();
    for (int i = threadIdx.x; i < MG_CI * xw; i += MG_THREADS) {
      const int ci = ci0 + i / xw, tt = i - (i / xw) * xw;
      float v = 0.f;
      if (ci < cin_total) {
        const int t = reflect(t0 - halo + tt, a.T);
        if (t >= 0 && t < a.T) {
          if (ci < a.cin) {
            v = __ldg(a.x + (static_cast<size_t>(b) * a.cin + ci) * a.T + t);
            if (a.lrelu_x) v = lrelu02(v);
          } else {
            v = __ldg(a.x2 + (static_cast<size_t>(b) * a.cin2 + (ci - a.cin)) * a.T + t);
            if (a.lrelu_x2) v = lrelu02(v);
          }
        }
      }
      xs[i] = v;
    }
    for (int i = threadIdx.x; i < MG_CI * KS * MG_CO; i += MG_THREADS) {
      const int co = i % MG_CO, ck = i / MG_CO;        // ck = ci * KS + k
      const int ci = ci0 + ck / KS;
      ws[i] = (ci < cin_total && co0 + co < a.cout) ? __ldg(a.w + (static_cast<size_t>(ci0) * KS + ck) * a.cout + co0 + co) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < MG_CI; ++ci) {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const float4 w4 = *reinterpret_cast<const float4*>(ws + (ci * KS + k) * MG_CO + ty * 4);
        const float* xr = xs + ci * xw + k * a.dil + tx;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xv = xr[16 * i];
          acc[0][i] = fmaf(w4.x, xv, acc[0][i]);
          acc[1][i] = fmaf(w4.y, xv, acc[1][i]);
          acc[2][i] = fmaf(w4.z, xv, acc[2][i]);
          acc[3][i] = fmaf(w4.w, xv, acc[3][i]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int co = co0 + ty * 4 + c;
    if (co >= a.cout) continue;
    const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
    float* o = a.out + (static_cast<size_t>(b) * a.cout + co) * a.T;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 + tx + 16 * i;
      if (t < a.T) o[t] = acc[c][i] + bv;
    }
  }
}

struct ConvTArgs {
  const float* x; int cin; int Tin;   // (B, cin, Tin), LeakyReLU applied on load
  const float* w;                     // packed [cin][K = 2 * stride][cout]
  const float* bias;
  float* out; int cout; int stride; int pad;
};

// out[t] = bias + sum_ci sum_{j in {0,1}} lrelu(x[ci][i - j]) * w[ci][k0 + j * s][co],  i = (t + pad) / s, k0 = (t + pad) % s
template <int S>
__global__ void __launch_bounds__(MG_THREADS)
convt1d_kernel(const ConvTArgs a) {
  extern __shared__ __align__(16) float mg_smem[];
  constexpr int K = 2 * S;
  constexpr int XW = MG_T / S + 2;                     // input samples a 128-step output tile touches (+1 earlier tap, +1 phase spill)
  float* xs = mg_smem;                                 // [MG_CI][XW]
  float* ws = mg_smem + MG_CI * XW;                    // [MG_CI][K][MG_CO]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int t0 = blockIdx.x * MG_T, co0 = blockIdx.y * MG_CO, b = blockIdx.z;
  const int Tout = a.Tin * S;
  const int i_base = (t0 + a.pad) / S - 1;             // first staged input sample
  const int tp = t0 + tx + a.pad;
  const int k0 = tp % S;                               // tap phase of all 8 outputs of this thread (16 is a multiple of S)
  const int il = tp / S - i_base;                      // staged index of the later tap of output 0
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
  for (int ci0 = 0; ci0 < a.cin; ci0 += MG_CI) {
    __syncthreads();
    for (int i = threadIdx.x; i < MG_CI * XW; i += MG_THREADS) {
      const int ci = ci0 + i / XW, ii = i_base + (i % XW);
      float v = 0.f;
      if (ci < a.cin && ii >= 0 && ii < a.Tin) v = lrelu02(__ldg(a.x + (static_cast<size_t>(b) * a.cin + ci) * a.Tin + ii));
      xs[i] = v;
    }
    for (int i = threadIdx.x; i < MG_CI * K * MG_CO; i += MG_THREADS) {
      const int co = i % MG_CO, ck = i / MG_CO;
      const int ci = ci0 + ck / K;
      ws[i] = (ci < a.cin && co0 + co < a.cout) ? __ldg(a.w + (static_cast<size_t>(ci0) * K + ck) * a.cout + co0 + co) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < MG_CI; ++ci) {
      const float4 wa = *reinterpret_cast<const float4*>(ws + (ci * K + k0) * MG_CO + ty * 4);       // tap k0     <- x[i]
      const float4 wb = *reinterpret_cast<const float4*>(ws + (ci * K + k0 + S) * MG_CO + ty * 4);   // tap k0 + s <- x[i - 1]
      const float* xr = xs + ci * XW + il;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xa = xr[(16 / S) * i], xb = xr[(16 / S) * i - 1];
        acc[0][i] = fmaf(wa.x, xa, fmaf(wb.x, xb, acc[0][i]));
        acc[1][i] = fmaf(wa.y, xa, fmaf(wb.y, xb, acc[1][i]));
        acc[2][i] = fmaf(wa.z, xa, fmaf(wb.z, xb, acc[2][i]));
        acc[3][i] = fmaf(wa.w, xa, fmaf(wb.w, xb, acc[3][i]));
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int co = co0 + ty * 4 + c;
    if (co >= a.cout) continue;
    const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
    float* o = a.out + (static_cast<size_t>(b) * a.cout + co) * Tout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 + tx + 16 * i;
      if (t < Tout) o[t] = acc[c][i] + bv;
    }
  }
}

// wave[b][t] = tanh(bias + sum_ci sum_k w[ci][k] * lrelu(x[b][ci][reflect(t + k - 3)]))
__global__ void __launch_bounds__(256)
conv_out_kernel(const float* __restrict__ x, int cin, int T, const float* __restrict__ w /* [cin][7] */, const float* __restrict__ bias,
                float* __restrict__ out) {
  extern __shared__ float mg_smem[];
  for (int i = threadIdx.x; i < cin * 7; i += blockDim.x) mg_smem[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int idx[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) idx[k] = reflect(t + k - 3, T);
  float s = bias ? bias[0] : 0.f;
  for (int ci = 0; ci < cin; ++ci) {
    const float* xr = x + (static_cast<size_t>(b) * cin + ci) * T;
#pragma unroll
    for (int k = 0; k < 7; ++k) s = fmaf(mg_smem[ci * 7 + k], lrelu02(__ldg(xr + idx[k])), s);
  }
  out[static_cast<size_t>(b) * T + t] = tanhf(s);
}

template <int KS>
int launch_conv(const ConvArgs& a, int B, cudaStream_t s) {
  const int halo = (KS - 1) / 2 * a.dil;
  const size_t smem = (static_cast<size_t>(MG_CI) * (MG_T + 2 * halo) + MG_CI * KS * MG_CO) * sizeof(float);
  dim3 grid(ceil_div(a.T, MG_T), ceil_div(a.cout, MG_CO), B);
  conv1d_kernel<KS><<<grid, MG_THREADS, smem, s>>>(a);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

template <int S>
int launch_convt(const ConvTArgs& a, int B, cudaStream_t s) {
  const size_t smem = (static_cast<size_t>(MG_CI) * (MG_T / S + 2) + MG_CI * 2 * S * MG_CO) * sizeof(float);
  dim3 grid(ceil_div(a.Tin * S, MG_T), ceil_div(a.cout, MG_CO), B);
  convt1d_kernel<S><<<grid, MG_THREADS, smem, s>>>(a);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

}  // namespace

// One packed convolution: w [cin_total][K][cout], bias [cout]
struct MgConv {
  float* w = nullptr;
  float* bias = nullptr;
  int cin = 0, cout = 0, k = 0;
  bool loaded_w = false, loaded_b = false;
};

struct Melgan {
  int n_mel = 0, ngf = 0, n_res = 0;
  MgConv conv_in;                       // model.1
  MgConv up[4];                         // model.{3,8,13,18} for three residual layers
  std::vector<MgConv> res_c3[4];        // block.2 of every ResnetBlock
  std::vector<MgConv> res_pair[4];      // [shortcut ; block.4] stacked along the input channels, biases summed
  MgConv conv_out;                      // model.24
  float* buf[3] = {nullptr, nullptr, nullptr};
  size_t buf_elems = 0;
  int launches = 0;
};

namespace {
const int kRatios[4] = {8, 8, 2, 2};

int alloc_conv(MgConv& c, int cin, int cout, int k) {
  c.cin = cin; c.cout = cout; c.k = k;
  MGV_CHECK_CUDA(cudaMalloc(&c.w, static_cast<size_t>(cin) * k * cout * sizeof(float)));
  MGV_CHECK_CUDA(cudaMalloc(&c.bias, cout * sizeof(float)));
  MGV_CHECK_CUDA(cudaMemset(c.bias, 0, cout * sizeof(float)));
  return MGV_OK;
}

// src (cout, cin, k) [conv] or (cin, cout, k) [transposed] fp32 -> dst [cin_off + cin][k][cout]
__global__ void pack_conv_weight_kernel(const float* __restrict__ src, int cout, int cin, int k, int transposed, int cin_off,
                                        float* __restrict__ dst) {
  const int n = cout * cin * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int co, ci, kk;
    if (transposed) { ci = i / (cout * k); co = (i / k) % cout; kk = i % k; }
    else { co = i / (cin * k); ci = (i / k) % cin; kk = i % k; }
    dst[(static_cast<size_t>(cin_off + ci) * k + kk) * cout + co] = src[i];
  }
}
__global__ void add_bias_kernel(const float* __restrict__ src, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
}  // namespace

int melgan_create(Melgan** out, int n_mel, int ngf, int n_res) {
  MGV_REQUIRE(out && n_mel >= 1 && ngf >= 1 && n_res >= 0 && n_res <= 8, "melgan_create: bad arguments");
  MGV_TRY(check_device());
  Melgan* m = new Melgan();
  m->n_mel = n_mel; m->ngf = ngf; m->n_res = n_res;
  int mult = 16;
  MGV_TRY(alloc_conv(m->conv_in, n_mel, mult * ngf, 7));
  for (int i = 0; i < 4; ++i) {
    const int cin = mult * ngf, cout = cin / 2;
    MGV_TRY(alloc_conv(m->up[i], cin, cout, 2 * kRatios[i]));
    m->res_c3[i].resize(n_res);
    m->res_pair[i].resize(n_res);
    for (int j = 0; j < n_res; ++j) {
      MGV_TRY(alloc_conv(m->res_c3[i][j], cout, cout, 3));
      MGV_TRY(alloc_conv(m->res_pair[i][j], 2 * cout, cout, 1));
    }
    mult /= 2;
  }
  MGV_TRY(alloc_conv(m->conv_out, ngf, 1, 7));
  *out = m;
  return MGV_OK;
}

int melgan_destroy(Melgan* m) {
  if (!m) return MGV_OK;
  auto fr = [](MgConv& c) { cudaFree(c.w); cudaFree(c.bias); };
  fr(m->conv_in); fr(m->conv_out);
  for (int i = 0; i < 4; ++i) {
    fr(m->up[i]);
    for (auto& c : m->res_c3[i]) fr(c);
    for (auto& c : m->res_pair[i]) fr(c);
  }
  for (float* b : m->buf) cudaFree(b);
  delete m;
  return MGV_OK;
}

// name = key of the reference Generator's state_dict with the weight norm already applied by the caller:
//   "model.<i>.weight" (effective weight g * v / |v|), "model.<i>.bias", "model.<i>.block.2.weight", "model.<i>.block.4.weight",
//   "model.<i>.shortcut.weight" and their biases  (vocoder/modules.py:46-77: layer indices of the nn.Sequential)
int melgan_load_weight(Melgan* m, const char* name, const float* src, long long numel, cudaStream_t s) {
  MGV_REQUIRE(m && name && src, "melgan_load_weight: null");
  int idx = -1;
  char rest[64] = {0};
  MGV_REQUIRE(sscanf(name, "model.%d.%63s", &idx, rest) == 2, "melgan_load_weight: unknown tensor '%s'", name);
  const std::string r(rest);
  const int per_stage = 2 + m->n_res;            // LeakyReLU, ConvTranspose1d, n_res ResnetBlocks
  const int last = 2 + 4 * per_stage + 2;        // index of the output convolution
  MgConv* c = nullptr;
  int cin_off = 0, transposed = 0;
  bool is_bias = false, pair_bias = false;
  auto weight_or_bias = [&](const std::string& tail) -> bool {
    if (tail == "weight") return true;
    if (tail == "bias") { is_bias = true; return true; }
    return false;
  };
  if (idx == 1) {
    c = &m->conv_in;
    MGV_REQUIRE(weight_or_bias(r), "melgan_load_weight: unknown tensor '%s'", name);
  } else if (idx == last) {
    c = &m->conv_out;
    MGV_REQUIRE(weight_or_bias(r), "melgan_load_weight: unknown tensor '%s'", name);
  } else {
    MGV_REQUIRE(idx >= 2 && idx < 2 + 4 * per_stage, "melgan_load_weight: layer index in '%s' out of range", name);
    const int st = (idx - 2) / per_stage, pos = (idx - 2) % per_stage;
    MGV_REQUIRE(pos >= 1, "melgan_load_weight: '%s' names a LeakyReLU", name);
    if (pos == 1) {
      c = &m->up[st];
      transposed = 1;
      MGV_REQUIRE(weight_or_bias(r), "melgan_load_weight: unknown tensor '%s'", name);
    } else {
      const int j = pos - 2;
      if (r.rfind("block.2.", 0) == 0) {
        c = &m->res_c3[st][j];
        MGV_REQUIRE(weight_or_bias(r.substr(8)), "melgan_load_weight: unknown tensor '%s'", name);
      } else if (r.rfind("shortcut.", 0) == 0) {
        c = &m->res_pair[st][j];
        cin_off = 0;
        pair_bias = true;
        MGV_REQUIRE(weight_or_bias(r.substr(9)), "melgan_load_weight: unknown tensor '%s'", name);
      } else if (r.rfind("block.4.", 0) == 0) {
        c = &m->res_pair[st][j];
        cin_off = c->cout;
        pair_bias = true;
        MGV_REQUIRE(weight_or_bias(r.substr(8)), "melgan_load_weight: unknown tensor '%s'", name);
      } else {
        MGV_REQUIRE(false, "melgan_load_weight: unknown tensor '%s'", name);
      }
    }
  }
  if (is_bias) {
    MGV_REQUIRE(numel == c->cout, "melgan_load_weight(%s): numel %lld != %d", name, numel, c->cout);
    if (pair_bias) {
      // the pair's bias is the sum of both convolutions' biases: loading a state_dict loads each exactly once after a reset
      add_bias_kernel<<<ceil_div(c->cout, 256), 256, 0, s>>>(src, c->cout, c->bias);
      MGV_CHECK_CUDA(cudaGetLastError());
    } else {
      MGV_CHECK_CUDA(cudaMemcpyAsync(c->bias, src, numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return MGV_OK;
  }
  const int cin = pair_bias ? c->cout : c->cin;
  MGV_REQUIRE(numel == static_cast<long long>(cin) * c->cout * c->k, "melgan_load_weight(%s): numel %lld != %d x %d x %d", name, numel,
              c->cout, cin, c->k);
  pack_conv_weight_kernel<<<ceil_div(static_cast<int>(numel), 256), 256, 0, s>>>(src, c->cout, cin, c->k, transposed, cin_off, c->w);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

// zero the summed pair biases before a state_dict is (re)loaded
int melgan_reset_biases(Melgan* m, cudaStream_t s) {
  MGV_REQUIRE(m, "melgan_reset_biases: null");
  for (int i = 0; i < 4; ++i)
    for (auto& c : m->res_pair[i]) MGV_CHECK_CUDA(cudaMemsetAsync(c.bias, 0, c.cout * sizeof(float), s));
  return MGV_OK;
}

// mel (B, n_mel, T) fp32 -> wave (B, 1, 256 T) fp32     (Generator.forward, vocoder/modules.py:79-80)
int melgan_forward(Melgan* m, const float* mel, int B, int T, float* wave, cudaStream_t s) {
  MGV_REQUIRE(m && B >= 0 && T >= 4, "melgan_forward: B=%d T=%d (the reflection padding needs T >= 4)", B, T);
  if (B == 0) return MGV_OK;
  MGV_REQUIRE(mel && wave, "melgan_forward: null pointer");
  // largest activation: the last three stages hold (4, 2, 1) ngf channels at (64, 128, 256) T steps = 256 ngf T per clip
  const size_t need = static_cast<size_t>(B) * 256 * m->ngf * T + 64;
  if (need > m->buf_elems) {
    for (float*& b : m->buf) { cudaFree(b); b = nullptr; }
    for (float*& b : m->buf) MGV_CHECK_CUDA(cudaMalloc(&b, need * sizeof(float)));
    m->buf_elems = need;
  }
  m->launches = 0;
  float* cur = m->buf[0];
  float* tmp = m->buf[1];
  float* nxt = m->buf[2];
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = mel; a.cin = m->n_mel; a.w = m->conv_in.w; a.bias = m->conv_in.bias; a.out = cur; a.cout = m->conv_in.cout; a.T = T; a.dil = 1;
  MGV_TRY(launch_conv<7>(a, B, s));
  m->launches++;
  int Tc = T;
  for (int i = 0; i < 4; ++i) {
    ConvTArgs u;
    u.x = cur; u.cin = m->up[i].cin; u.Tin = Tc; u.w = m->up[i].w; u.bias = m->up[i].bias; u.out = nxt; u.cout = m->up[i].cout;
    u.stride = kRatios[i]; u.pad = kRatios[i] / 2 + kRatios[i] % 2;
    if (kRatios[i] == 8) MGV_TRY(launch_convt<8>(u, B, s)); else MGV_TRY(launch_convt<2>(u, B, s));
    m->launches++;
    Tc *= kRatios[i];
    std::swap(cur, nxt);
    const int C = m->up[i].cout;
    int dil = 1;
    for (int j = 0; j < m->n_res; ++j, dil *= 3) {
      MGV_REQUIRE(dil < Tc, "melgan_forward: dilation %d needs more than %d steps", dil, Tc);
      ConvArgs c3;
      memset(&c3, 0, sizeof(c3));
      c3.x = cur; c3.cin = C; c3.lrelu_x = 1; c3.w = m->res_c3[i][j].w; c3.bias = m->res_c3[i][j].bias; c3.out = tmp; c3.cout = C; c3.T = Tc;
      c3.dil = dil;
      MGV_TRY(launch_conv<3>(c3, B, s));
      ConvArgs p;
      memset(&p, 0, sizeof(p));
      p.x = cur; p.cin = C; p.lrelu_x = 0; p.x2 = tmp; p.cin2 = C; p.lrelu_x2 = 1; p.w = m->res_pair[i][j].w; p.bias = m->res_pair[i][j].bias;
      p.out = nxt; p.cout = C; p.T = Tc; p.dil = 1;
      MGV_TRY(launch_conv<1>(p, B, s));
      m->launches += 2;
      std::swap(cur, nxt);
    }
  }
  dim3 grid(ceil_div(Tc, 256), B);
  conv_out_kernel<<<grid, 256, m->ngf * 7 * sizeof(float), s>>>(cur, m->ngf, Tc, m->conv_out.w, m->conv_out.bias, wave);
  MGV_CHECK_CUDA(cudaGetLastError());
  m->launches++;
  return MGV_OK;
}

int melgan_last_launches(const Melgan* m) { return m ? m->launches : 0; }

}  // namespace mgv
