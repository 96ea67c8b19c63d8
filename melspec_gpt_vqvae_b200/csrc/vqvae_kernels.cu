// Non-GEMM kernels of the VQVAE encoder / decoder.  See vqvae_kernels.cuh.
#include "vqvae_kernels.cuh"

namespace mgv {

namespace {

constexpr float GN_EPS = 1e-6f;   // Normalize(): GroupNorm(32, C, eps=1e-6)  (big_model_attn_gan.py:139-140)
constexpr int GN_GROUPS = 32;

inline int grid_for(long long items, int threads, int per_sm = 8) {
  long long b = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ------------------------------------------------------------------ weight prep
__global__ void repack_conv_kernel(const float* __restrict__ src, int Cout, int Cin, int KK,
                                   __nv_bfloat16* __restrict__ dst) {
  const long long total = static_cast<long long>(Cout) * Cin * KK;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    const long long r = i / Cin;
    const int tap = static_cast<int>(r % KK);
    const int co = static_cast<int>(r / KK);
    dst[i] = __float2bfloat16(src[(static_cast<long long>(co) * Cin + ci) * KK + tap]);
  }
}

__global__ void gather_table_kernel(const float* __restrict__ codebook, const float* __restrict__ wpq,
                                    const float* __restrict__ bpq, int K, int Cin, int Cout,
                                    __nv_bfloat16* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * Cout) return;
  const int co = i % Cout, k = i / Cout;
  float acc = 0.f;
  for (int c = 0; c < Cin; ++c) acc = fmaf(codebook[static_cast<size_t>(k) * Cin + c], wpq[static_cast<size_t>(co) * Cin + c], acc);
  table[i] = __float2bfloat16(acc + bpq[co]);
}

__global__ void gather_rows_kernel(const long long* __restrict__ idx, const uint4* __restrict__ table, long long n,
                                   int C8, int K, uint4* __restrict__ out, int* __restrict__ bad_flag) {
  const long long total = n * C8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / C8;
    const int c8 = static_cast<int>(i - row * C8);
    long long code = idx[row];
    if (code < 0 || code >= K) {
      if (bad_flag) atomicExch(bad_flag, 1);
      code = 0;
    }
    out[i] = __ldg(table + code * C8 + c8);
  }
}

// ------------------------------------------------------------------ layout changes
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ in, int N, int C, int HW,
                                         __nv_bfloat16* __restrict__ out) {
  const long long total = static_cast<long long>(N) * C * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const int p = static_cast<int>(r % HW);
    const long long n = r / HW;
    out[i] = __float2bfloat16(in[(n * C + c) * HW + p]);
  }
}

__global__ void nhwc_to_nchw_f32_kernel(const float* __restrict__ in, int N, int C, int HW, float* __restrict__ out) {
  const long long total = static_cast<long long>(N) * C * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % HW);
    const long long r = i / HW;
    const int c = static_cast<int>(r % C);
    const long long n = r / C;
    out[i] = in[(n * HW + p) * C + c];
  }
}

// ------------------------------------------------------------------ GroupNorm
// grid (chunks, N); block 256; every thread owns a fixed channel octet (256 % (C/8) == 0)
__global__ void __launch_bounds__(256)
gn_stats_kernel(const uint4* __restrict__ x, int HW, int C, float* __restrict__ part) {
  __shared__ float4 sh[256];   // per-thread (sum0, sq0, sum1, sq1) of its two 4-channel halves
  const int C8 = C / 8;
  const int n = blockIdx.y;
  const long long items = static_cast<long long>(HW) * C8;
  const uint4* xn = x + static_cast<long long>(n) * items;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < items;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(xn + i);
    float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    s0 += (a.x + a.y) + (b.x + b.y);
    q0 += (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y);
    s1 += (c.x + c.y) + (d.x + d.y);
    q1 += (c.x * c.x + c.y * c.y) + (d.x * d.x + d.y * d.y);
  }
  sh[threadIdx.x] = make_float4(s0, q0, s1, q1);
  __syncthreads();
  // deterministic fold: thread (g, stat) walks the contributing threads in a fixed order (no atomics)
  if (threadIdx.x < GN_GROUPS * 2) {
    const int g = threadIdx.x >> 1, stat = threadIdx.x & 1;
    const int gch = C / GN_GROUPS;  // 4, 8, 16 channels per group
    float acc = 0.f;
    for (int t = 0; t < 256; ++t) {
      const int c0 = (t % C8) * 8;
      const float4 v = sh[t];
      if (c0 / gch == g) acc += stat ? v.y : v.x;
      if ((c0 + 4) / gch == g) acc += stat ? v.w : v.z;
    }
    part[(static_cast<long long>(n) * gridDim.x + blockIdx.x) * GN_GROUPS * 2 + threadIdx.x] = acc;
  }
}

// one warp per (image, group): lanes walk the tiles in a fixed order, then a shuffle tree -> deterministic
__global__ void gn_finalize_kernel(const float* __restrict__ part, int total, int tiles, float inv_cnt,
                                   float* __restrict__ mr) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= total) return;
  const int n = w / GN_GROUPS, g = w - n * GN_GROUPS;
  const float2* p = reinterpret_cast<const float2*>(part) + static_cast<long long>(n) * tiles * GN_GROUPS + g;
  float s = 0.f, q = 0.f;
  for (int t = lane; t < tiles; t += 32) {
    const float2 v = p[static_cast<long long>(t) * GN_GROUPS];
    s += v.x;
    q += v.y;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    const float mean = s * inv_cnt;
    const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
    mr[2 * w] = mean;
    mr[2 * w + 1] = rsqrtf(var + GN_EPS);
  }
}

__device__ __forceinline__ void gn_mean_rstd(const float* __restrict__ mr, long long n, int g, float /*inv_cnt*/,
                                             float& mean, float& rstd) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(mr) + n * GN_GROUPS + g);
  mean = v.x;
  rstd = v.y;
}

__device__ __forceinline__ uint4 gn_apply8(const uint4 v, float m0, float r0, float m1, float r1,
                                            const float* __restrict__ gamma, const float* __restrict__ beta, int c0,
                                            bool do_swish) {
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0));
  const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
  const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + c0));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  float o[8];
  o[0] = (a.x - m0) * r0 * ga.x + ba.x;
  o[1] = (a.y - m0) * r0 * ga.y + ba.y;
  o[2] = (b.x - m0) * r0 * ga.z + ba.z;
  o[3] = (b.y - m0) * r0 * ga.w + ba.w;
  o[4] = (c.x - m1) * r1 * gb.x + bb.x;
  o[5] = (c.y - m1) * r1 * gb.y + bb.y;
  o[6] = (d.x - m1) * r1 * gb.z + bb.z;
  o[7] = (d.y - m1) * r1 * gb.w + bb.w;
  if (do_swish) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = swish(o[e]);
  }
  return make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}

// grid (chunks, N); block 256; every thread owns a fixed channel octet (256 % (C/8) == 0), so the group
// statistics and the affine parameters are loop invariant and the loop body is load / 8 FMA (+swish) / store.
__global__ void __launch_bounds__(256)
gn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ mr, const float* __restrict__ gamma,
                const float* __restrict__ beta, int HW, int C, int do_swish, uint4* __restrict__ y) {
  const int C8 = C / 8;
  const int gch = C / GN_GROUPS;
  const int n = blockIdx.y;
  const int c8 = threadIdx.x % C8;
  const int c0 = c8 * 8;
  float m0, r0, m1, r1;
  gn_mean_rstd(mr, n, c0 / gch, 0.f, m0, r0);
  gn_mean_rstd(mr, n, (c0 + 4) / gch, 0.f, m1, r1);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0));
  const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
  const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + c0));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
  // y = x * sc + sh
  const float sc[8] = {r0 * ga.x, r0 * ga.y, r0 * ga.z, r0 * ga.w, r1 * gb.x, r1 * gb.y, r1 * gb.z, r1 * gb.w};
  const float sh[8] = {ba.x - m0 * sc[0], ba.y - m0 * sc[1], ba.z - m0 * sc[2], ba.w - m0 * sc[3],
                       bb.x - m1 * sc[4], bb.y - m1 * sc[5], bb.z - m1 * sc[6], bb.w - m1 * sc[7]};
  const long long items = static_cast<long long>(HW) * C8;
  const uint4* xn = x + static_cast<long long>(n) * items;
  uint4* yn = y + static_cast<long long>(n) * items;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < items; i += 2 * stride) {
    const bool two = i + stride < items;
    const uint4 v0 = __ldg(xn + i);
    uint4 v1 = make_uint4(0, 0, 0, 0);
    if (two) v1 = __ldg(xn + i + stride);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      const uint4 v = h ? v1 : v0;
      const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
      float o[8] = {fmaf(a.x, sc[0], sh[0]), fmaf(a.y, sc[1], sh[1]), fmaf(b.x, sc[2], sh[2]), fmaf(b.y, sc[3], sh[3]),
                    fmaf(c.x, sc[4], sh[4]), fmaf(c.y, sc[5], sh[5]), fmaf(d.x, sc[6], sh[6]), fmaf(d.y, sc[7], sh[7])};
      if (do_swish) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = swish(o[e]);
      }
      yn[i + h * stride] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                      pack_bf16x2(o[6], o[7]));
    }
  }
}

// ------------------------------------------------------------------ nearest 2x upsample
__global__ void upsample2x_kernel(const uint4* __restrict__ x, int N, int H, int W, int C8, uint4* __restrict__ y) {
  const long long total = static_cast<long long>(N) * (2 * H) * (2 * W) * C8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    long long r = i / C8;
    const int xo = static_cast<int>(r % (2 * W));
    r /= (2 * W);
    const int yo = static_cast<int>(r % (2 * H));
    const long long n = r / (2 * H);
    y[i] = __ldg(x + ((n * H + (yo >> 1)) * W + (xo >> 1)) * C8 + c8);
  }
}

// ------------------------------------------------------------------ spatial self-attention (AttnBlock :434-446)
// Single-head attention over the T = 5 x 53 latent positions with C = 512 channels, on the legacy tensor path
// (mma.sync m16n8k16 bf16 -> fp32).  CTA = 32 query rows of one image, 8 warps.  Three phases over shared memory:
//   scores   per 32-key chunk every warp owns one 16 x 8 tile of S = Q K^T (32 k-steps over the 512 channels),
//   softmax  exact, over the full fp32 score rows (32 x T floats stay in shared memory),
//   P V      per 32-key chunk every warp owns a 16-row x 128-channel slab of O (16 accumulator tiles), P converted
//            to bf16 fragments on the fly, V fragments through ldmatrix.trans.
// Rows are padded by 16 bytes (stride 1040 B): fragment loads and ldmatrix rows hit distinct banks.
// (First version: the same structure on CUDA-core FMAs, 560 us per launch at 64 images.)
constexpr int SA_ROWS = 32;
constexpr int SA_KCH = 32;
constexpr int SA_THREADS = 256;

__device__ __forceinline__ void sa_mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void sa_ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r0), "=r"(r1)
               : "r"(smem_u32(smem_row)));
}

template <int C>
__global__ void __launch_bounds__(SA_THREADS)
spatial_attn_kernel(const __nv_bfloat16* __restrict__ qkv, int T, __nv_bfloat16* __restrict__ o) {
  constexpr int RS = C / 2 + 4;            // row stride in 32-bit words (16-byte aligned rows, conflict-free fragments)
  extern __shared__ __align__(16) uint32_t sa_smem[];
  uint32_t* Qs = sa_smem;                  // [32][RS]  bf16 pairs
  uint32_t* KVs = Qs + SA_ROWS * RS;       // [32][RS]
  const int Tpad = ((T + SA_KCH - 1) / SA_KCH) * SA_KCH;
  const int SW = Tpad + 1;
  float* Ss = reinterpret_cast<float*>(KVs + SA_KCH * RS);  // [32][SW]

  const int n = blockIdx.y;
  const int q0 = blockIdx.x * SA_ROWS;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int g = lane >> 2, tq = lane & 3;  // mma fragment coordinates
  const __nv_bfloat16* base = qkv + static_cast<long long>(n) * T * (3 * C);
  constexpr int V8 = C / 8;                // uint4 per row

  // ---- Q tile
  for (int i = t; i < SA_ROWS * V8; i += SA_THREADS) {
    const int r = i / V8, c8 = i - r * V8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < T) v = *reinterpret_cast<const uint4*>(base + static_cast<long long>(q0 + r) * (3 * C) + c8 * 8);
    *reinterpret_cast<uint4*>(Qs + r * RS + c8 * 4) = v;
  }
  const float scale = rsqrtf(static_cast<float>(C));   // int(c) ** (-0.5)   (:439)
  const int nchunks = Tpad / SA_KCH;
  auto load_chunk = [&](int kc, int col_off) {   // K (col_off = C) or V (2C) rows of chunk kc -> KVs, zero beyond T
    for (int i = t; i < SA_KCH * V8; i += SA_THREADS) {
      const int r = i / V8, c8 = i - r * V8;
      const int key = kc * SA_KCH + r;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (key < T) v = *reinterpret_cast<const uint4*>(base + static_cast<long long>(key) * (3 * C) + col_off + c8 * 8);
      *reinterpret_cast<uint4*>(KVs + r * RS + c8 * 4) = v;
    }
  };

  // ---- scores S = Q K^T * scale: warp -> query rows 16*(warp & 1) .. +15, keys 8*(warp >> 1) .. +7 of the chunk
  {
    const int mrow = (warp & 1) * 16, ncol = (warp >> 1) * 8;
    for (int kc = 0; kc < nchunks; ++kc) {
      __syncthreads();
      load_chunk(kc, C);
      __syncthreads();
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t* qa = Qs + (mrow + g) * RS + tq;
      const uint32_t* kb = KVs + (ncol + g) * RS + tq;
#pragma unroll 8
      for (int ks = 0; ks < C / 16; ++ks) {          // 16 channels = 8 words per step
        uint32_t a[4];
        a[0] = qa[ks * 8];
        a[1] = qa[8 * RS + ks * 8];
        a[2] = qa[ks * 8 + 4];
        a[3] = qa[8 * RS + ks * 8 + 4];
        sa_mma_16816(acc, a, kb[ks * 8], kb[ks * 8 + 4]);
      }
      float* s0 = Ss + (mrow + g) * SW + kc * SA_KCH + ncol + tq * 2;
      s0[0] = acc[0] * scale;
      s0[1] = acc[1] * scale;
      s0[8 * SW] = acc[2] * scale;
      s0[8 * SW + 1] = acc[3] * scale;
    }
  }
  __syncthreads();

  // ---- softmax over keys (dim=2, :440)
  for (int r = warp; r < SA_ROWS; r += SA_THREADS / 32) {
    float* row = Ss + r * SW;
    float mx = -INFINITY;
    for (int k = lane; k < T; k += 32) mx = fmaxf(mx, row[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < Tpad; k += 32) {
      const float e = (k < T) ? expf(row[k] - mx) : 0.f;
      row[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int k = lane; k < Tpad; k += 32) row[k] *= inv;
  }

  // ---- O = P V: warp -> query rows 16*(warp & 1) .. +15, channels 128*(warp >> 1) .. +127 (16 tiles of 8)
  const int mrow = (warp & 1) * 16, cbase = (warp >> 1) * (C / 4);
  float oacc[C / 32][4];
#pragma unroll
  for (int i = 0; i < C / 32; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  for (int kc = 0; kc < nchunks; ++kc) {
    __syncthreads();
    load_chunk(kc, 2 * C);
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < SA_KCH / 16; ++ks) {       // 16 keys per step
      const float* p0 = Ss + (mrow + g) * SW + kc * SA_KCH + ks * 16 + tq * 2;
      uint32_t a[4];
      a[0] = pack_bf16x2(p0[0], p0[1]);
      a[1] = pack_bf16x2(p0[8 * SW], p0[8 * SW + 1]);
      a[2] = pack_bf16x2(p0[8], p0[9]);
      a[3] = pack_bf16x2(p0[8 * SW + 8], p0[8 * SW + 9]);
      const uint32_t* vrow = KVs + (ks * 16 + (lane & 15)) * RS + cbase / 2;
#pragma unroll
      for (int i = 0; i < C / 32; ++i) {
        uint32_t b0, b1;
        sa_ldmatrix_x2_trans(b0, b1, vrow + i * 4);
        sa_mma_16816(oacc[i], a, b0, b1);
      }
    }
  }
  const int r0 = q0 + mrow + g, r1 = r0 + 8;
#pragma unroll
  for (int i = 0; i < C / 32; ++i) {
    const int ch = cbase + i * 8 + tq * 2;
    if (r0 < T)
      *reinterpret_cast<uint32_t*>(o + (static_cast<long long>(n) * T + r0) * C + ch) = pack_bf16x2(oacc[i][0], oacc[i][1]);
    if (r1 < T)
      *reinterpret_cast<uint32_t*>(o + (static_cast<long long>(n) * T + r1) * C + ch) = pack_bf16x2(oacc[i][2], oacc[i][3]);
  }
}

// ------------------------------------------------------------------ decoder tail: norm_out + swish + conv_out (C -> 1)
// One warp = CO_R output rows x a strip of columns of one image.  Lane l owns channels 4l..4l+3 (exactly one
// GroupNorm group at C = 128): its 9 x 4 filter weights and its normalisation constants live in registers, and it
// slides a 3-column window of ACTIVATED inputs (CO_R + 2 rows) along x -- every input pixel is loaded once per row
// group as one coalesced 256-byte row and activated once, the next column's loads are in flight while the current
// one is reduced.  Per column: 36 FMAs per output row and lane, then a 6-shuffle transposing reduction over the
// lanes.  (First version: shared-memory tile + per-tap shared-memory weight reads, 2.84 ms for 64 clips.)
constexpr int CO_R = 4;        // output rows per warp
constexpr int CO_STRIPS = 4;   // column strips per row group
constexpr int CO_THREADS = 128;

template <int C>
__global__ void __launch_bounds__(CO_THREADS)
norm_swish_conv_out_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ sums,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ w, const float* __restrict__ bias, int N, int H, int W,
                           float* __restrict__ out) {
  static_assert(C == 128, "lane <-> 4 channels <-> one GroupNorm group");
  const int lane = threadIdx.x & 31;
  const int rgs = (H + CO_R - 1) / CO_R;
  const int sw = (W + CO_STRIPS - 1) / CO_STRIPS;
  long long wid = static_cast<long long>(blockIdx.x) * (CO_THREADS / 32) + (threadIdx.x >> 5);
  if (wid >= static_cast<long long>(N) * rgs * CO_STRIPS) return;
  const int strip = static_cast<int>(wid % CO_STRIPS);
  wid /= CO_STRIPS;
  const int rg = static_cast<int>(wid % rgs);
  const int n = static_cast<int>(wid / rgs);
  const int y0 = rg * CO_R;
  const int xs = strip * sw, xe = min(W, xs + sw);
  if (xs >= xe) return;
  const int c0 = lane * 4;

  float mean, rstd;
  gn_mean_rstd(sums, n, lane, 0.f, mean, rstd);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c0));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0));
  float wt[9][4];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(w + tap * C + c0));
    wt[tap][0] = q.x; wt[tap][1] = q.y; wt[tap][2] = q.z; wt[tap][3] = q.w;
  }
  const float bval = __ldg(bias);

  // raw 8-byte loads of column x for rows y0-1 .. y0+CO_R (zero = outside the image: padding applies AFTER activation)
  auto load_col = [&](int x, uint2 (&raw)[CO_R + 2], unsigned& valid) {
    valid = 0;
#pragma unroll
    for (int r = 0; r < CO_R + 2; ++r) {
      const int y = y0 - 1 + r;
      raw[r] = make_uint2(0u, 0u);
      if (x >= 0 && x < W && y >= 0 && y < H) {
        raw[r] = __ldg(reinterpret_cast<const uint2*>(h + ((static_cast<long long>(n) * H + y) * W + x) * C + c0));
        valid |= 1u << r;
      }
    }
  };
  auto activate = [&](const uint2 (&raw)[CO_R + 2], unsigned valid, float (&a)[CO_R + 2][4]) {
#pragma unroll
    for (int r = 0; r < CO_R + 2; ++r) {
      const float2 p = unpack_bf16x2(raw[r].x), q = unpack_bf16x2(raw[r].y);
      float o[4];
      o[0] = (p.x - mean) * rstd * ga.x + be.x;
      o[1] = (p.y - mean) * rstd * ga.y + be.y;
      o[2] = (q.x - mean) * rstd * ga.z + be.z;
      o[3] = (q.y - mean) * rstd * ga.w + be.w;
      const bool ok = (valid >> r) & 1u;
#pragma unroll
      for (int e = 0; e < 4; ++e)   // fp32 activation (never stored)
        a[r][e] = ok ? swish(o[e]) : 0.f;
    }
  };

  float a0[CO_R + 2][4], a1[CO_R + 2][4], a2[CO_R + 2][4];   // columns x-1, x, x+1
  uint2 raw[CO_R + 2];
  unsigned valid;
  load_col(xs - 1, raw, valid);
  activate(raw, valid, a0);
  load_col(xs, raw, valid);
  activate(raw, valid, a1);
  load_col(xs + 1, raw, valid);
  for (int x = xs; x < xe; ++x) {
    activate(raw, valid, a2);
    load_col(x + 2, raw, valid);          // in flight during the reduction below
    float acc[CO_R];
#pragma unroll
    for (int r = 0; r < CO_R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s = fmaf(a0[r + dy][e], wt[dy * 3 + 0][e], s);
          s = fmaf(a1[r + dy][e], wt[dy * 3 + 1][e], s);
          s = fmaf(a2[r + dy][e], wt[dy * 3 + 2][e], s);
        }
      acc[r] = s;
    }
    // transposing reduction over the 32 lanes: 4 values -> lanes 0, 8, 16, 24 hold rows 0, 1, 2, 3
    const bool up16 = (lane & 16) != 0;
    float t0 = (up16 ? acc[2] : acc[0]) + __shfl_xor_sync(0xffffffffu, up16 ? acc[0] : acc[2], 16);
    float t1 = (up16 ? acc[3] : acc[1]) + __shfl_xor_sync(0xffffffffu, up16 ? acc[1] : acc[3], 16);
    const bool up8 = (lane & 8) != 0;
    float u = (up8 ? t1 : t0) + __shfl_xor_sync(0xffffffffu, up8 ? t0 : t1, 8);
    u += __shfl_xor_sync(0xffffffffu, u, 4);
    u += __shfl_xor_sync(0xffffffffu, u, 2);
    u += __shfl_xor_sync(0xffffffffu, u, 1);
    if ((lane & 7) == 0) {
      const int y = y0 + (lane >> 3);
      if (y < H) out[(static_cast<long long>(n) * H + y) * W + x] = u + bval;
    }
#pragma unroll
    for (int r = 0; r < CO_R + 2; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        a0[r][e] = a1[r][e];
        a1[r][e] = a2[r][e];
      }
  }
}

// ------------------------------------------------------------------ encoder head: conv_in 1 -> Cout
__global__ void __launch_bounds__(256)
conv_in_1ch_kernel(const float* __restrict__ mel, const float* __restrict__ w, const float* __restrict__ bias, int N,
                   int H, int W, int Cout, uint4* __restrict__ out) {
  // A thread keeps the 9 x 8 weights of its 8 output channels in registers and walks over pixels; the C8 = Cout/8
  // threads of a pixel are consecutive, so every store instruction writes whole pixels (Cout * 2 contiguous bytes).
  // Pixel coordinates advance by carries, not divisions.  (First version: weights re-read from shared memory per tap
  // and three 64-bit divisions per output vector -- 2.2 ms for 64 clips, 13 x the 0.17 ms its 1.1 GB of stores need.)
  const int C8 = Cout / 8;
  const int c8 = threadIdx.x % C8;                 // host guarantees blockDim % C8 == 0
  float wr[9][8], bs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    bs[e] = __ldg(bias + c8 * 8 + e);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wr[tap][e] = __ldg(w + (c8 * 8 + e) * 9 + tap);
  }
  const int ppb = blockDim.x / C8;                 // pixels per block and iteration
  const long long n_pix = static_cast<long long>(N) * H * W;
  const long long step = static_cast<long long>(gridDim.x) * ppb;
  long long p = static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / C8;
  if (p >= n_pix) return;
  int x = static_cast<int>(p % W);
  long long r = p / W;
  int y = static_cast<int>(r % H);
  int n = static_cast<int>(r / H);
  const int step_x = static_cast<int>(step % W);
  const long long step_r = step / W;
  const int step_y = static_cast<int>(step_r % H), step_n = static_cast<int>(step_r / H);
  for (; p < n_pix; p += step) {
    const float* base = mel + (static_cast<long long>(n) * H + y) * W + x;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bs[e];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = y + dy - 1, xx = x + dx - 1;
        const float m = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(base + (dy - 1) * W + (dx - 1)) : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(m, wr[dy * 3 + dx][e], acc[e]);
      }
    out[p * C8 + c8] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                                  pack_bf16x2(acc[6], acc[7]));
    // advance (x, y, n) by `step` pixels
    x += step_x;
    int cy = step_y;
    if (x >= W) {
      x -= W;
      ++cy;
    }
    y += cy;
    int cn = step_n;
    if (y >= H) {
      y -= H;
      ++cn;
    }
    n += cn;
  }
}

// Upsample (nearest 2x, then 3x3 conv pad 1: big_model_attn_gan.py:182-186) as four 2x2 convolutions over the LOW-res
// tensor: output pixel (2y+py, 2x+px) only ever sees input rows {y-1, y} (py = 0) or {y, y+1} (py = 1), likewise in x, so
// the filter rows / columns that land on the same input pixel are summed once here (fp32, then one rounding to bf16).
// src fp32 OIHW (Cout, Cin, 3, 3) -> dst bf16 [phase = py*2+px][Cout][tap = ty*2+tx][Cin]
__global__ void upsample_phase_weights_kernel(const float* __restrict__ src, int Cout, int Cin, __nv_bfloat16* __restrict__ dst) {
  const long long total = 16LL * Cout * Cin;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cin);
    long long r = i / Cin;
    const int tap = static_cast<int>(r & 3);
    r >>= 2;
    const int o = static_cast<int>(r % Cout);
    const int ph = static_cast<int>(r / Cout);
    const int py = ph >> 1, px = ph & 1, ty = tap >> 1, tx = tap & 1;
    // filter rows that fall on input row (y + ty - (1 - py)): py=0: {0} | {1,2};  py=1: {0,1} | {2}
    const int ky0 = (py == 0) ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2);
    const int ky1 = (py == 0) ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
    const int kx0 = (px == 0) ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2);
    const int kx1 = (px == 0) ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
    const float* w = src + (static_cast<long long>(o) * Cin + c) * 9;
    float acc = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) acc += w[ky * 3 + kx];
    dst[i] = __float2bfloat16(acc);
  }
}

}  // namespace

// ====================================================================== host wrappers
int vqvae_upsample_phase_weights(const float* src, int Cout, int Cin, __nv_bfloat16* dst, cudaStream_t s) {
  upsample_phase_weights_kernel<<<grid_for(16LL * Cout * Cin, 256), 256, 0, s>>>(src, Cout, Cin, dst);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_repack_conv_weight(const float* src, int Cout, int Cin, int KH, int KW, __nv_bfloat16* dst, cudaStream_t s) {
  const long long total = static_cast<long long>(Cout) * Cin * KH * KW;
  repack_conv_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, Cout, Cin, KH * KW, dst);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_build_gather_table(const float* codebook, const float* wpq, const float* bpq, int K, int Cin, int Cout,
                             __nv_bfloat16* table, cudaStream_t s) {
  gather_table_kernel<<<ceil_div(K * Cout, 256), 256, 0, s>>>(codebook, wpq, bpq, K, Cin, Cout, table);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_gather_rows(const long long* idx, const __nv_bfloat16* table, long long n, int C, int K, __nv_bfloat16* out,
                      int* bad_flag, cudaStream_t s) {
  MGV_REQUIRE(C % 8 == 0, "gather_rows: C=%d", C);
  if (n == 0) return MGV_OK;
  gather_rows_kernel<<<grid_for(n * (C / 8), 256), 256, 0, s>>>(idx, reinterpret_cast<const uint4*>(table), n, C / 8, K,
                                                                reinterpret_cast<uint4*>(out), bad_flag);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_nchw_f32_to_nhwc_bf16(const float* in, int N, int C, int HW, __nv_bfloat16* out, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * C * HW;
  if (total == 0) return MGV_OK;
  nchw_to_nhwc_bf16_kernel<<<grid_for(total, 256), 256, 0, s>>>(in, N, C, HW, out);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_nhwc_f32_to_nchw_f32(const float* in, int N, int C, int HW, float* out, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * C * HW;
  if (total == 0) return MGV_OK;
  nhwc_to_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, s>>>(in, N, C, HW, out);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_gn_stats(const __nv_bfloat16* x, int N, int HW, int C, float* part, int* tiles_out, cudaStream_t s) {
  MGV_REQUIRE(C % GN_GROUPS == 0 && (C == 128 || C == 256 || C == 512), "gn_stats: C=%d unsupported", C);
  const long long items = static_cast<long long>(HW) * (C / 8);
  int chunks = static_cast<int>((items + 256 * 8 - 1) / (256 * 8));
  const int cap = (num_sms() * 8 + (N > 0 ? N : 1) - 1) / (N > 0 ? N : 1);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  if (tiles_out) *tiles_out = chunks;
  if (N == 0) return MGV_OK;
  gn_stats_kernel<<<dim3(chunks, N), 256, 0, s>>>(reinterpret_cast<const uint4*>(x), HW, C, part);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_gn_finalize(const float* part, int N, int tiles, float count, float* mr, cudaStream_t s) {
  if (N == 0) return MGV_OK;
  const int total = N * GN_GROUPS;   // warps
  gn_finalize_kernel<<<ceil_div(total * 32, 256), 256, 0, s>>>(part, total, tiles, 1.0f / count, mr);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_gn_apply(const __nv_bfloat16* x, const float* mr, const float* gamma, const float* beta, int N, int HW, int C,
                   int do_swish, __nv_bfloat16* y, cudaStream_t s) {
  MGV_REQUIRE(C == 128 || C == 256 || C == 512, "gn_apply: C=%d unsupported", C);
  if (N == 0 || HW == 0) return MGV_OK;
  const long long items = static_cast<long long>(HW) * (C / 8);
  int chunks = static_cast<int>((items + 256 * 4 - 1) / (256 * 4));
  const int cap = (num_sms() * 16 + N - 1) / N;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  gn_apply_kernel<<<dim3(chunks, N), 256, 0, s>>>(reinterpret_cast<const uint4*>(x), mr, gamma, beta, HW, C, do_swish,
                                                  reinterpret_cast<uint4*>(y));
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_upsample2x(const __nv_bfloat16* x, int N, int H, int W, int C, __nv_bfloat16* y, cudaStream_t s) {
  MGV_REQUIRE(C % 8 == 0, "upsample: C=%d", C);
  const long long total = static_cast<long long>(N) * 4 * H * W * (C / 8);
  if (total == 0) return MGV_OK;
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(x), N, H, W, C / 8,
                                                         reinterpret_cast<uint4*>(y));
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_spatial_attention(const __nv_bfloat16* qkv, int N, int T, int C, __nv_bfloat16* o, cudaStream_t s) {
  MGV_REQUIRE(C == 512, "spatial_attention: C=%d unsupported (the reference attends only at the 512-channel level)", C);
  MGV_REQUIRE(T >= 1 && T <= 1024, "spatial_attention: T=%d", T);
  if (N == 0) return MGV_OK;
  constexpr int RS = 512 / 2 + 4;
  const int Tpad = ceil_div(T, SA_KCH) * SA_KCH;
  const size_t smem = (static_cast<size_t>(SA_ROWS) * RS + SA_KCH * RS) * 4 + static_cast<size_t>(SA_ROWS) * (Tpad + 1) * 4;
  static unsigned long long attr_mask = 0;   // per device (and per template instantiation)
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(spatial_attn_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(spatial_attn_kernel<512>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
  }
  MGV_REQUIRE(smem <= 200 * 1024, "spatial_attention: T=%d needs too much shared memory", T);
  spatial_attn_kernel<512><<<dim3(ceil_div(T, SA_ROWS), N), SA_THREADS, smem, s>>>(qkv, T, o);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_norm_swish_conv_out(const __nv_bfloat16* h, const float* sums, const float* gamma, const float* beta,
                              const float* w, const float* bias, int N, int H, int W, int C, float* out,
                              cudaStream_t s) {
  MGV_REQUIRE(C == 128, "conv_out: C=%d unsupported", C);
  if (N == 0) return MGV_OK;
  const long long warps = static_cast<long long>(N) * ceil_div(H, CO_R) * CO_STRIPS;
  const long long blocks = (warps + CO_THREADS / 32 - 1) / (CO_THREADS / 32);
  norm_swish_conv_out_kernel<128><<<static_cast<unsigned>(blocks), CO_THREADS, 0, s>>>(h, sums, gamma, beta, w, bias, N, H, W, out);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vqvae_conv_in_1ch(const float* mel, const float* w, const float* bias, int N, int H, int W, int Cout,
                      __nv_bfloat16* out, cudaStream_t s) {
  MGV_REQUIRE(Cout % 8 == 0 && Cout <= 512 && 256 % (Cout / 8) == 0, "conv_in: Cout=%d", Cout);
  const long long n_pix = static_cast<long long>(N) * H * W;
  if (n_pix == 0) return MGV_OK;
  const int ppb = 256 / (Cout / 8);
  long long blocks = (n_pix + ppb - 1) / ppb;
  const long long cap = static_cast<long long>(num_sms()) * 16;   // ~8 pixels per thread: the weight preload amortises
  if (blocks > cap) blocks = cap;
  conv_in_1ch_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(mel, w, bias, N, H, W, Cout, reinterpret_cast<uint4*>(out));
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

}  // namespace mgv
