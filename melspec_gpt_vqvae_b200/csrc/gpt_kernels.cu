// Non-GEMM kernels of the minGPT path.  See gpt_kernels.cuh.
#include <stdlib.h>
#include "gpt_kernels.cuh"

namespace mgv {

namespace {

// ------------------------------------------------------------------ embedding
__global__ void embed_kernel(const long long* __restrict__ idx, int R, int p_off, int idx_ld,
                             const float* __restrict__ prefix_emb, const long long* __restrict__ cls,
                             const float* __restrict__ embedder, int m, const float* __restrict__ tok_emb,
                             const float* __restrict__ pos_emb, int C, int vocab, int class_size,
                             float* __restrict__ x_out, int* __restrict__ err_flag) {
  pdl_launch_dependents();   // dependents may start their prologue / weight prefetch now
  pdl_wait();                // ... but our inputs need the upstream grid
  const int row = blockIdx.x;  // b*R + r
  const int b = row / R, p = p_off + (row - b * R);
  const float* src;
  if (p < m) {
    if (prefix_emb) {
      src = prefix_emb + (static_cast<long long>(b) * m + p) * C;
    } else {
      long long c = cls[b];
      if (c < 0 || c >= class_size) {
        if (err_flag) atomicExch(err_flag, 1);
        c = 0;
      }
      src = embedder + c * C;
    }
  } else {
    long long tok = idx[static_cast<long long>(b) * idx_ld + (p - m)];
    if (tok < 0 || tok >= vocab) {
      if (err_flag) atomicExch(err_flag, 2);
      tok = 0;
    }
    src = tok_emb + tok * C;
  }
  const float4* s4 = reinterpret_cast<const float4*>(src);
  const float4* p4 = reinterpret_cast<const float4*>(pos_emb + static_cast<long long>(p) * C);
  float4* o4 = reinterpret_cast<float4*>(x_out + static_cast<long long>(row) * C);
  for (int i = threadIdx.x; i < C / 4; i += blockDim.x) {
    const float4 a = __ldg(s4 + i), q = __ldg(p4 + i);
    o4[i] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  }
}

// ------------------------------------------------------------------ LayerNorm
constexpr int LN_MAX_V4 = 16;  // C <= 2048

__global__ void __launch_bounds__(128)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bia, int rows, int C,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ zero_buf, long long zero_count) {
  pdl_launch_dependents();   // dependents may start their prologue / weight prefetch now
  pdl_wait();                // ... but our inputs need the upstream grid
  (void)zero_buf;
  (void)zero_count;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= rows) return;
  const int nv = C / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  float4 v[LN_MAX_V4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_V4; ++j) {
    const int i = lane + 32 * j;
    if (i < nv) {
      v[j] = x4[i];
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
  // parameters: issue the loads now so that they are in flight during the two reductions (decode: C = 1024 -> 8 + 8
  // float4 per lane; the latency of a dependent load is a measurable part of a 3 us stage)
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(bia);
  float4 gq[LN_MAX_V4 / 2], bq[LN_MAX_V4 / 2];
#pragma unroll
  for (int j = 0; j < LN_MAX_V4 / 2; ++j) {
    const int i = lane + 32 * j;
    if (i < nv) {
      gq[j] = __ldg(w4 + i);
      bq[j] = __ldg(b4 + i);
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(C);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_V4; ++j) {
    const int i = lane + 32 * j;
    if (i < nv) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(C) + 1e-5f);
  uint2* o2 = reinterpret_cast<uint2*>(out + static_cast<long long>(row) * C);
#pragma unroll
  for (int j = 0; j < LN_MAX_V4; ++j) {
    const int i = lane + 32 * j;
    if (i < nv) {
      const float4 g = (j < LN_MAX_V4 / 2) ? gq[j] : __ldg(w4 + i), be = (j < LN_MAX_V4 / 2) ? bq[j] : __ldg(b4 + i);
      const float a = (v[j].x - mean) * rstd * g.x + be.x;
      const float b = (v[j].y - mean) * rstd * g.y + be.y;
      const float c = (v[j].z - mean) * rstd * g.z + be.z;
      const float d = (v[j].w - mean) * rstd * g.w + be.w;
      o2[i] = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
    }
  }
}

// Throughput form for the prefill / teacher-forced forward (thousands of rows): the decode form above keeps 16 float4 of x
// and 16 float4 of parameters per lane for C <= 2048 (~170 registers, 12 rows in flight per SM: 3.0 TB/s measured at
// 16 960 x 1024, 34.9 us).  Here the row length is a template parameter and the parameters are read where they are used
// (they stay in L1), so a lane holds NV float4 only and an SM keeps 5 x more rows in flight: 16.9 us = 6.2 TB/s.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bia, int rows,
                      __nv_bfloat16* __restrict__ out) {
  constexpr int C = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float4* x4 = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = x4[lane + 32 * j];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / C) + 1e-5f);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(bia);
  uint2* o2 = reinterpret_cast<uint2*>(out + static_cast<long long>(row) * C);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    const float4 g = __ldg(w4 + i), be = __ldg(b4 + i);
    const float a = (v[j].x - mean) * rstd * g.x + be.x;
    const float b = (v[j].y - mean) * rstd * g.y + be.y;
    const float c = (v[j].z - mean) * rstd * g.z + be.z;
    const float d = (v[j].w - mean) * rstd * g.w + be.w;
    o2[i] = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
  }
}

// ------------------------------------------------------------------ prefill attention
// Causal (or prefix-unmasked) attention over a whole sequence, T <= 288, head dim 64, on the legacy tensor path
// (mma.sync m16n8k16 bf16 -> fp32; this part is 4% of the prefill FLOPs, the GEMMs around it are tcgen05).
// One CTA = one (sequence, head): K and V of the head are staged ONCE in shared memory (rows padded to 144 B:
// conflict-free fragment loads and ldmatrix) and 9 warps walk the 16-row query blocks in causal-balanced pairs
// (block w and block n-1-w see the same number of keys in total).  Flash-style single pass with an online softmax
// (running max / sum per row, accumulator rescaled per 64-key block) -- except when the caller wants the attention
// map (the last layer): then pass 1 finds the exact row maximum and sum and pass 2 emits the normalised
// probabilities (fp32) while it accumulates P V.  No [T,T] tensor reaches HBM otherwise.
// (First version: one CTA per 64 query rows, K/V restaged per CTA, always two passes: 370 us per layer at
// bs=64 against 100 us for all four GEMMs of the layer.)
constexpr int FA_BN = 64;          // keys per block
constexpr int FA_LD = 72;          // bf16 elements per smem row (64 + 8 pad)
constexpr int FA_WARPS = 9;        // ceil(288 / 16 / 2) query-block pairs
constexpr int FA_THREADS = FA_WARPS * 32;
constexpr int FA_MAXK = 320;       // >= GPT_MAX_T rounded up to FA_BN

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r0), "=r"(r1)
               : "r"(smem_u32(smem_row)));
}

template <bool WRITE_ATT>
__global__ void __launch_bounds__(FA_THREADS, 2)   // two CTAs per SM (92 KB of smem each): <= 113 registers
attn_prefill_kernel(const __nv_bfloat16* __restrict__ qkv, int B, int T, int nh, int n_unmasked,
                    __nv_bfloat16* __restrict__ y, float* __restrict__ att, int att_T,
                    __nv_bfloat16* __restrict__ kcache, __nv_bfloat16* __restrict__ vcache, int Tmax) {
  extern __shared__ __align__(16) __nv_bfloat16 fa_smem[];
  const int kpad = ((T + FA_BN - 1) / FA_BN) * FA_BN;
  __nv_bfloat16* sK = fa_smem;                       // [kpad][72]
  __nv_bfloat16* sV = sK + kpad * FA_LD;             // [kpad][72]
  const int bh = blockIdx.x;
  const int b = bh / nh, h = bh - b * nh;
  const int C = nh * GPT_HEAD_DIM;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  // ---- stage K and V (zero beyond T); also fills the KV cache
  for (int i = t; i < kpad * 8; i += FA_THREADS) {
    const int r = i >> 3, part = i & 7;
    uint4 kq = make_uint4(0, 0, 0, 0), vq = make_uint4(0, 0, 0, 0);
    if (r < T) {
      const __nv_bfloat16* base = qkv + (static_cast<long long>(b) * T + r) * (3 * C) + h * GPT_HEAD_DIM + part * 8;
      kq = *reinterpret_cast<const uint4*>(base + C);
      vq = *reinterpret_cast<const uint4*>(base + 2 * C);
      if (kcache != nullptr) {
        const long long co = ((static_cast<long long>(b) * nh + h) * Tmax + r) * GPT_HEAD_DIM + part * 8;
        *reinterpret_cast<uint4*>(kcache + co) = kq;
        *reinterpret_cast<uint4*>(vcache + co) = vq;
      }
    }
    *reinterpret_cast<uint4*>(sK + r * FA_LD + part * 8) = kq;
    *reinterpret_cast<uint4*>(sV + r * FA_LD + part * 8) = vq;
  }
  __syncthreads();

  const int nrb = (T + 15) / 16;                       // 16-row query blocks
  const int g = lane >> 2, tq = lane & 3;
  // 1/sqrt(d) (minGPT.py:81) folded with log2(e): probabilities are exp2(s' - max')
  const float scale2 = 1.4426950408889634f / sqrtf(static_cast<float>(GPT_HEAD_DIM));

  for (int turn = 0; turn < 2; ++turn) {
    // causal-balanced pairing: warp w takes block w, then block nrb-1-w
    const int rb = (turn == 0) ? warp : nrb - 1 - warp;
    if (turn == 0 ? (2 * warp >= nrb) : (rb <= warp)) continue;   // warp-uniform
    const int rbase = rb * 16;
    const int row0 = min(rbase + g, T - 1), row1 = min(rbase + g + 8, T - 1);   // padded rows mirror the last row
    const bool r0_ok = rbase + g < T, r1_ok = rbase + g + 8 < T;
    // Q fragments straight from global memory (each row block is read once)
    uint32_t aq[4][4];
    {
      const __nv_bfloat16* q0p = qkv + (static_cast<long long>(b) * T + row0) * (3 * C) + h * GPT_HEAD_DIM + tq * 2;
      const __nv_bfloat16* q1p = qkv + (static_cast<long long>(b) * T + row1) * (3 * C) + h * GPT_HEAD_DIM + tq * 2;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        aq[kk][0] = *reinterpret_cast<const uint32_t*>(q0p + kk * 16);
        aq[kk][1] = *reinterpret_cast<const uint32_t*>(q1p + kk * 16);
        aq[kk][2] = *reinterpret_cast<const uint32_t*>(q0p + kk * 16 + 8);
        aq[kk][3] = *reinterpret_cast<const uint32_t*>(q1p + kk * 16 + 8);
      }
    }
    // keys this block's rows can see: causal limit of its last row, or the unmasked prefix
    const int wlast = min(rbase + 15, T - 1);
    const int wkmax = (rbase < n_unmasked) ? max(wlast + 1, min(n_unmasked, T)) : wlast + 1;
    const int wnkb = (wkmax + FA_BN - 1) / FA_BN;

    auto scores = [&](int jb, float (&sc)[8][4]) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        sc[nb][0] = sc[nb][1] = sc[nb][2] = sc[nb][3] = 0.f;
        const __nv_bfloat16* kb = sK + (jb * FA_BN + nb * 8 + g) * FA_LD + tq * 2;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kb + kk * 16);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kb + kk * 16 + 8);
          mma_bf16_16816(sc[nb], aq[kk], b0, b1);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = jb * FA_BN + nb * 8 + tq * 2 + (e & 1);
          const int row = (e < 2) ? row0 : row1;
          // mask[i][j] = tril, plus the unmasked prefix block (minGPT.py:65-68, :82)
          const bool allowed = key < T && (key <= row || (row < n_unmasked && key < n_unmasked));
          sc[nb][e] = allowed ? sc[nb][e] * scale2 : -INFINITY;
        }
      }
    };
    auto block_max = [&](const float (&sc)[8][4], float& bm0, float& bm1) {
      bm0 = -INFINITY;
      bm1 = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        bm0 = fmaxf(bm0, fmaxf(sc[nb][0], sc[nb][1]));
        bm1 = fmaxf(bm1, fmaxf(sc[nb][2], sc[nb][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    };
    auto pv = [&](int jb, const float (&sc)[8][4], float (&o)[8][4]) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {     // 16 keys per step
        uint32_t ap[4];
        ap[0] = pack_bf16x2(sc[2 * ks][0], sc[2 * ks][1]);
        ap[1] = pack_bf16x2(sc[2 * ks][2], sc[2 * ks][3]);
        ap[2] = pack_bf16x2(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
        ap[3] = pack_bf16x2(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
        const __nv_bfloat16* vrow = sV + (jb * FA_BN + ks * 16 + (lane & 15)) * FA_LD;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
          uint32_t b0, b1;
          ldmatrix_x2_trans(b0, b1, vrow + nd * 8);
          mma_bf16_16816(o[nd], ap, b0, b1);
        }
      }
    };

    float o[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // l: this lane's share of the row sum
    if (!WRITE_ATT) {
      // ---- single pass, online softmax (key 0 is always allowed -> the maximum is finite from block 0 on)
      for (int jb = 0; jb < wnkb; ++jb) {
        float sc[8][4];
        scores(jb, sc);
        float bm0, bm1;
        block_max(sc, bm0, bm1);
        const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
        const float a0 = exp2f(m0 - n0), a1 = exp2f(m1 - n1);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          sc[nb][0] = exp2f(sc[nb][0] - n0);
          sc[nb][1] = exp2f(sc[nb][1] - n0);
          sc[nb][2] = exp2f(sc[nb][2] - n1);
          sc[nb][3] = exp2f(sc[nb][3] - n1);
          s0 += sc[nb][0] + sc[nb][1];
          s1 += sc[nb][2] + sc[nb][3];
        }
        l0 = l0 * a0 + s0;
        l1 = l1 * a1 + s1;
        m0 = n0;
        m1 = n1;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
          o[nd][0] *= a0;
          o[nd][1] *= a0;
          o[nd][2] *= a1;
          o[nd][3] *= a1;
        }
        pv(jb, sc, o);
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        o[nd][0] *= inv0;
        o[nd][1] *= inv0;
        o[nd][2] *= inv1;
        o[nd][3] *= inv1;
      }
    } else {
      // ---- pass 1: exact row maximum and sum
      for (int jb = 0; jb < wnkb; ++jb) {
        float sc[8][4];
        scores(jb, sc);
        float bm0, bm1;
        block_max(sc, bm0, bm1);
        const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          s0 += exp2f(sc[nb][0] - n0) + exp2f(sc[nb][1] - n0);
          s1 += exp2f(sc[nb][2] - n1) + exp2f(sc[nb][3] - n1);
        }
        l0 = l0 * exp2f(m0 - n0) + s0;
        l1 = l1 * exp2f(m1 - n1) + s1;
        m0 = n0;
        m1 = n1;
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
      // ---- pass 2: normalised probabilities -> attention map, and O = P V
      float* arow0 = (att && r0_ok && row0 < att_T) ? att + ((static_cast<long long>(b) * nh + h) * att_T + row0) * att_T : nullptr;
      float* arow1 = (att && r1_ok && row1 < att_T) ? att + ((static_cast<long long>(b) * nh + h) * att_T + row1) * att_T : nullptr;
      for (int jb = 0; jb < wnkb; ++jb) {
        float sc[8][4];
        scores(jb, sc);
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          sc[nb][0] = exp2f(sc[nb][0] - m0) * inv0;
          sc[nb][1] = exp2f(sc[nb][1] - m0) * inv0;
          sc[nb][2] = exp2f(sc[nb][2] - m1) * inv1;
          sc[nb][3] = exp2f(sc[nb][3] - m1) * inv1;
          const int key = jb * FA_BN + nb * 8 + tq * 2;
          if (arow0) {   // buffer is pre-zeroed: write only non-zero probabilities
            if (key < att_T && sc[nb][0] != 0.f) arow0[key] = sc[nb][0];
            if (key + 1 < att_T && sc[nb][1] != 0.f) arow0[key + 1] = sc[nb][1];
          }
          if (arow1) {
            if (key < att_T && sc[nb][2] != 0.f) arow1[key] = sc[nb][2];
            if (key + 1 < att_T && sc[nb][3] != 0.f) arow1[key + 1] = sc[nb][3];
          }
        }
        pv(jb, sc, o);
      }
    }
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      const int dim = nd * 8 + tq * 2;
      if (r0_ok)
        *reinterpret_cast<uint32_t*>(y + (static_cast<long long>(b) * T + rbase + g) * C + h * GPT_HEAD_DIM + dim) =
            pack_bf16x2(o[nd][0], o[nd][1]);
      if (r1_ok)
        *reinterpret_cast<uint32_t*>(y + (static_cast<long long>(b) * T + rbase + g + 8) * C + h * GPT_HEAD_DIM + dim) =
            pack_bf16x2(o[nd][2], o[nd][3]);
    }
  }
}

// ------------------------------------------------------------------ decode attention
// One CTA per (sequence, head), 128 threads = 16 key groups x 8 dim chunks: thread (kg, dc) owns the 8
// head dims [8*dc, 8*dc+8) (one 16-byte load per key) of keys kg, kg+16, ... ; keys are walked in passes of
// 64 (4 loads per thread) with two register buffers, so the loads of pass p+1 are in flight while pass p is
// reduced: the cache streams continuously instead of in load / compute phases.  A warp's load covers 4
// consecutive 128-byte cache rows (512 contiguous bytes).  <= 72 registers and ~2 KB smem so that all
// B*n_head CTAs are resident in ONE wave (7 per SM at B=64): the step's latency is one dependency chain,
// not one per wave.  The first K pass is issued before the grid dependency resolves, the first V pass before
// the softmax reductions.
constexpr int AD_THREADS = 128;
constexpr int AD_KG = 16;                 // key groups
constexpr int AD_KPP = 4;                 // keys per thread per pass
constexpr int AD_PASS = AD_KG * AD_KPP;   // 64 keys per pass

__global__ void __launch_bounds__(AD_THREADS, 7)
attn_decode_kernel(float* __restrict__ qkv32, int nh, const int* __restrict__ pos_ptr,
                   __nv_bfloat16* __restrict__ kcache, __nv_bfloat16* __restrict__ vcache, int Tmax,
                   __nv_bfloat16* __restrict__ y, float* __restrict__ att_rows, int Tatt, int zero_consumed,
                   float* __restrict__ zero_buf, int zero_per_cta, const LnFold fold) {
  pdl_launch_dependents();
  __shared__ __align__(16) float sq[GPT_HEAD_DIM];
  __shared__ __align__(16) __nv_bfloat16 sk[GPT_HEAD_DIM];
  __shared__ __align__(16) __nv_bfloat16 sv[GPT_HEAD_DIM];
  __shared__ float ss[GPT_MAX_T];
  __shared__ float red[8];
  __shared__ float sacc[4][GPT_HEAD_DIM];
  const int bh = blockIdx.x;
  const int b = bh / nh, h = bh - b * nh;
  const int C = nh * GPT_HEAD_DIM;
  const int pos = *pos_ptr;   // written by an earlier decode step (not by the upstream grid)
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int kg = t >> 3, dc = t & 7;
  const __nv_bfloat16* kc = kcache + (static_cast<long long>(b) * nh + h) * Tmax * GPT_HEAD_DIM + dc * 8;
  const __nv_bfloat16* vc = vcache + (static_cast<long long>(b) * nh + h) * Tmax * GPT_HEAD_DIM + dc * 8;
  const int npass = pos / AD_PASS + 1;

  auto load_rows = [&](const __nv_bfloat16* base, const __nv_bfloat16* fresh, int p, uint4 (&buf)[AD_KPP]) {
#pragma unroll
    for (int i = 0; i < AD_KPP; ++i) {
      const int j = p * AD_PASS + kg + AD_KG * i;
      buf[i] = make_uint4(0, 0, 0, 0);
      if (j < pos) buf[i] = *reinterpret_cast<const uint4*>(base + static_cast<long long>(j) * GPT_HEAD_DIM);
      else if (j == pos && fresh) buf[i] = *reinterpret_cast<const uint4*>(fresh + dc * 8);
    }
  };

  // ---- first pass of K rows: positions < pos were written by earlier steps, so load before the dependency
  uint4 b0[AD_KPP], b1[AD_KPP];
  load_rows(kc, nullptr, 0, b0);
  // folded LayerNorm (gemm_decode_fold.cu): q, k, v arrive as raw accumulators of W (gamma.x); the fold vectors are
  // weights-like constants, so they are also fetched before the dependency resolves
  float f_sw[3] = {0.f, 0.f, 0.f}, f_bp[3] = {0.f, 0.f, 0.f};
  if (fold.stats != nullptr && t < GPT_HEAD_DIM) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      f_sw[i] = __ldg(fold.sw + i * C + h * GPT_HEAD_DIM + t);
      f_bp[i] = __ldg(fold.bp + i * C + h * GPT_HEAD_DIM + t);
    }
  }
  pdl_wait();
  // clear this CTA's slice of the FC1 split-K accumulator: its last reader (the previous layer's FC2) has
  // completed and the next writer (this layer's FC1) runs after this kernel
  if (zero_buf)
    for (int i = t; i < zero_per_cta; i += AD_THREADS) zero_buf[static_cast<long long>(blockIdx.x) * zero_per_cta + i] = 0.f;
  if (t < GPT_HEAD_DIM) {
    float* base = qkv32 + static_cast<long long>(b) * 3 * C + h * GPT_HEAD_DIM + t;
    float qf = __ldcg(base), kf = __ldcg(base + C), vf = __ldcg(base + 2 * C);
    if (fold.stats != nullptr) {
      float s = 0.f, ss = 0.f;
      for (int z = 0; z < fold.nparts; ++z) {     // fixed order: deterministic
        const float2 p = __ldcg(fold.stats + static_cast<long long>(z) * fold.stride + b);
        s += p.x;
        ss += p.y;
      }
      const float inv_n = 1.0f / static_cast<float>(fold.dim);
      const float mean = s * inv_n;
      const float rstd = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + 1e-5f);
      qf = fmaf(rstd, fmaf(-mean, f_sw[0], qf), f_bp[0]);
      kf = fmaf(rstd, fmaf(-mean, f_sw[1], kf), f_bp[1]);
      vf = fmaf(rstd, fmaf(-mean, f_sw[2], vf), f_bp[2]);
    }
    // q is rounded to bf16 like the prefill path (which stores q,k,v as bf16)
    sq[t] = __bfloat162float(__float2bfloat16(qf));
    const __nv_bfloat16 kb = __float2bfloat16(kf);
    const __nv_bfloat16 vb = __float2bfloat16(vf);
    if (zero_consumed) {  // the next layer's split-K QKV GEMM accumulates into this buffer with atomics
      base[0] = 0.f;
      base[C] = 0.f;
      base[2 * C] = 0.f;
    }
    sk[t] = kb;
    sv[t] = vb;
    const long long co = (static_cast<long long>(b) * nh + h) * Tmax * GPT_HEAD_DIM + static_cast<long long>(pos) * GPT_HEAD_DIM + t;
    kcache[co] = kb;
    vcache[co] = vb;
  }
  __syncthreads();

  // ---- scores
  const float scale = 1.0f / sqrtf(static_cast<float>(GPT_HEAD_DIM));
  float q8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) q8[e] = sq[dc * 8 + e];
  float lmax = -INFINITY;
  auto score_pass = [&](int p, const uint4 (&buf)[AD_KPP]) {
#pragma unroll
    for (int i = 0; i < AD_KPP; ++i) {
      const int j = p * AD_PASS + kg + AD_KG * i;
      uint4 kv = buf[i];
      if (j == pos) kv = *reinterpret_cast<const uint4*>(sk + dc * 8);
      const float2 a = unpack_bf16x2(kv.x), bq = unpack_bf16x2(kv.y), c = unpack_bf16x2(kv.z), d = unpack_bf16x2(kv.w);
      float s = q8[0] * a.x;
      s = fmaf(q8[1], a.y, s);
      s = fmaf(q8[2], bq.x, s);
      s = fmaf(q8[3], bq.y, s);
      s = fmaf(q8[4], c.x, s);
      s = fmaf(q8[5], c.y, s);
      s = fmaf(q8[6], d.x, s);
      s = fmaf(q8[7], d.y, s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s *= scale;
      if (j <= pos) {
        if (dc == 0) ss[j] = s;
        lmax = fmaxf(lmax, s);
      }
    }
  };
  for (int p = 0; p < npass; p += 2) {
    if (p + 1 < npass) load_rows(kc, nullptr, p + 1, b1);
    score_pass(p, b0);
    if (p + 2 < npass) load_rows(kc, nullptr, p + 2, b0);
    if (p + 1 < npass) score_pass(p + 1, b1);
  }
  // ---- V rows of the first pass: issue now, consume after the softmax reductions
  load_rows(vc, sv, 0, b0);
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  const float mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float lsum = 0.f;
  for (int j = t; j <= pos; j += AD_THREADS) {
    const float e = expf(ss[j] - mx);
    ss[j] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[4 + warp] = lsum;
  __syncthreads();
  const float inv = 1.0f / ((red[4] + red[5]) + (red[6] + red[7]));
  if (att_rows && pos < Tatt) {
    float* arow = att_rows + ((static_cast<long long>(b) * nh + h) * Tatt + pos) * Tatt;
    for (int j = t; j <= pos; j += AD_THREADS) arow[j] = ss[j] * inv;
  }
  // ---- PV
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto pv_pass = [&](int p, const uint4 (&buf)[AD_KPP]) {
#pragma unroll
    for (int i = 0; i < AD_KPP; ++i) {
      const int j = p * AD_PASS + kg + AD_KG * i;
      if (j <= pos) {
        const float pj = ss[j];
        const float2 a = unpack_bf16x2(buf[i].x), bq = unpack_bf16x2(buf[i].y), c = unpack_bf16x2(buf[i].z),
                     d = unpack_bf16x2(buf[i].w);
        acc[0] = fmaf(pj, a.x, acc[0]);
        acc[1] = fmaf(pj, a.y, acc[1]);
        acc[2] = fmaf(pj, bq.x, acc[2]);
        acc[3] = fmaf(pj, bq.y, acc[3]);
        acc[4] = fmaf(pj, c.x, acc[4]);
        acc[5] = fmaf(pj, c.y, acc[5]);
        acc[6] = fmaf(pj, d.x, acc[6]);
        acc[7] = fmaf(pj, d.y, acc[7]);
      }
    }
  };
  for (int p = 0; p < npass; p += 2) {
    if (p + 1 < npass) load_rows(vc, sv, p + 1, b1);
    pv_pass(p, b0);
    if (p + 2 < npass) load_rows(vc, sv, p + 2, b0);
    if (p + 1 < npass) pv_pass(p + 1, b1);
  }
  // reduce the 4 key groups of a warp (lanes with equal dc), then the 4 warps through smem
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) sacc[warp][lane * 8 + e] = acc[e];
  }
  __syncthreads();
  if (t < GPT_HEAD_DIM) {
    const float o = (sacc[0][t] + sacc[1][t]) + (sacc[2][t] + sacc[3][t]);
    y[static_cast<long long>(b) * C + h * GPT_HEAD_DIM + t] = __float2bfloat16(o * inv);
  }
}

// ------------------------------------------------------------------ GELU (split-K FC1 path)
constexpr int GELU_VEC = 1;   // float4 per thread (measured in one run at batch 64: 1 -> 873, 2 -> 881, 4 -> 897 us per position)
__global__ void __launch_bounds__(256)
gelu_bf16_kernel(float* __restrict__ h32, long long n4, __nv_bfloat16* __restrict__ out, int zero_consumed) {
  pdl_launch_dependents();   // dependents may start their prologue / weight prefetch now
  pdl_wait();                // ... but our inputs need the upstream grid
  float4* i4 = reinterpret_cast<float4*>(h32);
  uint2* o2 = reinterpret_cast<uint2*>(out);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < n4; i0 += GELU_VEC * stride) {
    float4 v[GELU_VEC];
#pragma unroll
    for (int q = 0; q < GELU_VEC; ++q)      // all loads first: independent L2 round trips
      if (i0 + q * stride < n4) v[q] = __ldcg(i4 + i0 + q * stride);
#pragma unroll
    for (int q = 0; q < GELU_VEC; ++q) {
      const long long i = i0 + q * stride;
      if (i < n4) {
        if (zero_consumed) i4[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // next layer's split-K FC1 accumulates here
        o2[i] = make_uint2(pack_bf16x2(gelu_erf(v[q].x), gelu_erf(v[q].y)), pack_bf16x2(gelu_erf(v[q].z), gelu_erf(v[q].w)));
      }
    }
  }
}

// ------------------------------------------------------------------ final LN + head + top-k + sample
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// Philox4x32-10 -> uniform in [0,1) with 24 bits
__device__ float philox_uniform(unsigned long long seed, uint32_t ctr0, uint32_t ctr1) {
  uint32_t c[4] = {ctr0, ctr1, 0u, 0u};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return static_cast<float>(c[0] >> 8) * (1.0f / 16777216.0f);
}

constexpr int SAMPLE_THREADS = 128;
constexpr int SAMPLE_MAX_V = 1024;

__device__ float block_reduce(float v, float* red, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < SAMPLE_THREADS / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// One CTA per sequence.  The logits of this position come from the head GEMM (ln_f and head run as the same
// LayerNorm / tcgen05 kernels as the blocks); here: temperature, top-k threshold (ties kept), softmax,
// multinomial / argmax, token append, embedding of the next position, position advance.
__global__ void __launch_bounds__(SAMPLE_THREADS)
sample_step_kernel(const SampleArgs a) {
  unsigned int* done_counter = a.done_counter;
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float sm_s[];
  float* sl = sm_s;            // [V]
  __shared__ float red[SAMPLE_THREADS / 32];
  __shared__ int s_tok;
  const int b = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31;
  const int pos = *a.pos_ptr;

  // ---- logits / temperature (minGPT.py:346); the accumulator is cleared for the next position's split-K head GEMM
  float* lrow = a.logits_acc + static_cast<long long>(b) * a.V;
  float f_mean = 0.f, f_rstd = 1.f;
  if (a.fold.stats != nullptr) {    // folded ln_f: raw accumulator of W_head (gamma.x) -> logits (gemm_decode_fold.cu)
    float s = 0.f, ss = 0.f;
    for (int z = 0; z < a.fold.nparts; ++z) {
      const float2 p = __ldcg(a.fold.stats + static_cast<long long>(z) * a.fold.stride + b);
      s += p.x;
      ss += p.y;
    }
    const float inv_n = 1.0f / static_cast<float>(a.fold.dim);
    f_mean = s * inv_n;
    f_rstd = rsqrtf(fmaxf(ss * inv_n - f_mean * f_mean, 0.f) + 1e-5f);
  }
  for (int i = t; i < a.V; i += SAMPLE_THREADS) {
    float raw = __ldcg(lrow + i);
    if (a.fold.stats != nullptr) raw = fmaf(f_rstd, fmaf(-f_mean, __ldg(a.fold.sw + i), raw), __ldg(a.fold.bp + i));
    const float l = __fdiv_rn(raw, a.temperature);
    lrow[i] = 0.f;
    sl[i] = l;
    if (a.logits_out) a.logits_out[static_cast<long long>(pos - a.logits_pos0) * a.logits_step_stride + static_cast<long long>(a.row0 + b) * a.V + i] = l;
  }
  __syncthreads();

  // ---- top_k_logits (minGPT.py:287-291): everything below the k-th largest value -> -inf (ties kept)
  if (a.top_k > 0 && a.top_k < a.V) {
    float cand = INFINITY;
    for (int i = t; i < a.V; i += SAMPLE_THREADS) {
      const float li = sl[i];
      int greater = 0;
      for (int j = 0; j < a.V; ++j) greater += (sl[j] > li) ? 1 : 0;
      if (greater < a.top_k) cand = fminf(cand, li);
    }
    const float thr = -block_reduce(-cand, red, true);
    __syncthreads();
    for (int i = t; i < a.V; i += SAMPLE_THREADS)
      if (sl[i] < thr) sl[i] = -INFINITY;
    __syncthreads();
  }

  // ---- softmax (:351)
  float lm = -INFINITY;
  for (int i = t; i < a.V; i += SAMPLE_THREADS) lm = fmaxf(lm, sl[i]);
  const float mx = block_reduce(lm, red, true);
  float ls = 0.f;
  for (int i = t; i < a.V; i += SAMPLE_THREADS) {
    const float e = (sl[i] == -INFINITY) ? 0.f : expf(sl[i] - mx);
    sl[i] = e;
    ls += e;
  }
  const float total = block_reduce(ls, red, false);
  __syncthreads();

  // ---- multinomial(probs, 1) (:354) or topk(probs, 1) (:356), by warp 0: every lane owns a contiguous slice
  if (t < 32) {
    const int per = (a.V + 31) / 32;
    const int lo = lane * per, hi = min(lo + per, a.V);
    int tok;
    if (a.do_sample) {
      float mine = 0.f;
      for (int i = lo; i < hi; ++i) mine += sl[i];
      float incl = mine;                       // inclusive scan over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const float u = philox_uniform(*a.seed_ptr, static_cast<uint32_t>(pos), static_cast<uint32_t>(a.row0 + b));
      const float target = u * __shfl_sync(0xffffffffu, incl, 31);
      const float excl = incl - mine;
      // first index whose cumulative probability exceeds the target
      int cand = 0x7fffffff;
      if (mine > 0.f && incl > target && excl <= target) {
        float cum = excl;
        for (int i = lo; i < hi; ++i) {
          cum += sl[i];
          if (sl[i] > 0.f && cum > target) {
            cand = i;
            break;
          }
        }
      }
      // fallbacks for rounding at the upper end: the last index with non-zero probability
      int last_nz = -1;
      for (int i = lo; i < hi; ++i)
        if (sl[i] > 0.f) last_nz = i;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        last_nz = max(last_nz, __shfl_xor_sync(0xffffffffu, last_nz, o));
      }
      tok = (cand != 0x7fffffff) ? cand : max(last_nz, 0);
    } else {
      float best = -1.f;
      int bi = 0x7fffffff;
      for (int i = lo; i < hi; ++i)
        if (sl[i] > best) {
          best = sl[i];
          bi = i;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      tok = bi;
    }
    (void)total;
    if (lane == 0) {
      s_tok = tok;
      const int slot = pos + 1 - a.m;
      if (slot >= 0 && slot < a.tokens_ld) a.tokens[static_cast<long long>(b) * a.tokens_ld + slot] = tok;
    }
  }
  __syncthreads();
  const int tok = s_tok;

  // ---- embedding of the sampled token for the next position (minGPT.py:170-180)
  if (a.x_next && pos + 1 < a.block_size) {
    const float4* te = reinterpret_cast<const float4*>(a.tok_emb + static_cast<long long>(tok) * a.C);
    const float4* pe = reinterpret_cast<const float4*>(a.pos_emb + static_cast<long long>(pos + 1) * a.C);
    float4* xn = reinterpret_cast<float4*>(a.x_next + static_cast<long long>(b) * a.C);
    for (int i = t; i < a.C / 4; i += SAMPLE_THREADS) {
      const float4 p = __ldg(te + i), q = __ldg(pe + i);
      xn[i] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
    }
  }

  // ---- the last CTA to finish advances the position (every CTA read *pos_ptr before arriving)
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(done_counter, 1u);
    if (done == gridDim.x - 1) {
      *a.pos_ptr = pos + 1;
      *done_counter = 0u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------ cross entropy per row
// loss[r] = logsumexp(logits[r, :]) - logits[r, target[r]]  (F.cross_entropy / nn.CrossEntropyLoss with unit class
// weights, reduction='none': transformer/minGPT.py:197, transformer/decoders.py:21,64-68).  One warp per row.
__global__ void __launch_bounds__(128)
ce_rows_kernel(const float* __restrict__ logits, const long long* __restrict__ targets, long long rows, int V,
               long long ld, float* __restrict__ loss, int* __restrict__ err_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (r >= rows) return;
  const float* row = logits + r * ld;
  float mx = -INFINITY;
  for (int i = lane; i < V; i += 32) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < V; i += 32) sum += expf(row[i] - mx);
  sum = warp_sum(sum);
  if (lane == 0) {
    const long long t = targets[r];
    if (t < 0 || t >= V) {
      *err_flag = 2;
      loss[r] = NAN;
    } else {
      loss[r] = (mx + logf(sum)) - row[t];
    }
  }
}

}  // namespace

int gpt_ce_rows(const float* logits, const long long* targets, long long rows, int V, long long ld, float* loss,
                int* err_flag, cudaStream_t s) {
  MGV_REQUIRE(logits && targets && loss && err_flag, "cross entropy: null");
  MGV_REQUIRE(V >= 1 && ld >= V, "cross entropy: V=%d ld=%lld", V, ld);
  if (rows == 0) return MGV_OK;
  const long long blocks = (rows + 3) / 4;
  MGV_REQUIRE(blocks <= 0x7fffffffLL, "cross entropy: too many rows");
  ce_rows_kernel<<<static_cast<unsigned>(blocks), 128, 0, s>>>(logits, targets, rows, V, ld, loss, err_flag);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

// ====================================================================== host wrappers
int gpt_embed(const long long* idx, int B, int R, int p_off, int idx_ld, const float* prefix_emb, const long long* cls,
              const float* embedder, int m, const float* tok_emb, const float* pos_emb, int C, int vocab,
              int class_size, float* x_out, int* err_flag, cudaStream_t s, bool pdl) {
  MGV_REQUIRE(C % 4 == 0, "embed: C=%d", C);
  MGV_REQUIRE(p_off >= m || prefix_emb || (cls && embedder), "embed: prefix of length %d without embeddings", m);
  MGV_REQUIRE(p_off + R <= m || idx, "embed: null idx");
  if (B * R == 0) return MGV_OK;
  LaunchCfg lc(dim3(B * R), dim3(256), 0, s, pdl);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, embed_kernel, idx, R, p_off, idx_ld, prefix_emb, cls, embedder, m, tok_emb,
                                    pos_emb, C, vocab, class_size, x_out, err_flag));
  return MGV_OK;
}

int gpt_layernorm(const float* x, const float* w, const float* b, int rows, int C, __nv_bfloat16* out, float* zero_buf,
                  long long zero_count, cudaStream_t s, bool pdl) {
  MGV_REQUIRE(C % 4 == 0 && C <= LN_MAX_V4 * 128, "layernorm: C=%d unsupported", C);
  MGV_REQUIRE(zero_count % 4 == 0, "layernorm: zero_count");
  if (rows == 0) return MGV_OK;
  if (rows >= 1024 && !pdl && zero_buf == nullptr && (C == 1024 || C == 512 || C == 256 || C == 2048)) {
    const int grid = ceil_div(rows, 8);
    switch (C) {
      case 256: layernorm_rows_kernel<2><<<grid, 256, 0, s>>>(x, w, b, rows, out); break;
      case 512: layernorm_rows_kernel<4><<<grid, 256, 0, s>>>(x, w, b, rows, out); break;
      case 1024: layernorm_rows_kernel<8><<<grid, 256, 0, s>>>(x, w, b, rows, out); break;
      default: layernorm_rows_kernel<16><<<grid, 256, 0, s>>>(x, w, b, rows, out); break;
    }
    MGV_CHECK_CUDA(cudaGetLastError());
    return MGV_OK;
  }
  LaunchCfg lc(dim3(ceil_div(rows, 4)), dim3(128), 0, s, pdl);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, layernorm_kernel, x, w, b, rows, C, out, zero_buf, zero_count));
  return MGV_OK;
}

int gpt_attention_prefill(const __nv_bfloat16* qkv, int B, int T, int nh, int n_unmasked, __nv_bfloat16* y, float* att,
                          int att_T, __nv_bfloat16* kcache, __nv_bfloat16* vcache, int Tmax, cudaStream_t s) {
  // plain causal mask (n_unmasked <= 1 is the tril mask), no attention map, no cache fill: the tcgen05 kernel
  if (att == nullptr && kcache == nullptr && n_unmasked <= 1 && T <= GPT_MAX_T && gpt_attention_prefill_tc_supported(T))
    return gpt_attention_prefill_tc(qkv, B, T, nh, y, s);
  return gpt_attention_prefill_mma(qkv, B, T, nh, n_unmasked, y, att, att_T, kcache, vcache, Tmax, s);
}

int gpt_attention_prefill_mma(const __nv_bfloat16* qkv, int B, int T, int nh, int n_unmasked, __nv_bfloat16* y, float* att,
                              int att_T, __nv_bfloat16* kcache, __nv_bfloat16* vcache, int Tmax, cudaStream_t s) {
  MGV_REQUIRE(T >= 1 && T <= GPT_MAX_T && (T + 15) / 16 <= 2 * FA_WARPS, "attention: T=%d exceeds %d", T, GPT_MAX_T);
  if (B == 0) return MGV_OK;
  const int kpad = ceil_div(T, FA_BN) * FA_BN;
  const size_t smem = static_cast<size_t>(2 * kpad) * FA_LD * sizeof(__nv_bfloat16);
  static unsigned long long attr_mask = 0;   // per device (and per template instantiation)
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        2 * FA_MAXK * FA_LD * 2));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        2 * FA_MAXK * FA_LD * 2));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
  }
  if (att != nullptr)
    attn_prefill_kernel<true><<<B * nh, FA_THREADS, smem, s>>>(qkv, B, T, nh, n_unmasked, y, att, att_T, kcache, vcache, Tmax);
  else
    attn_prefill_kernel<false><<<B * nh, FA_THREADS, smem, s>>>(qkv, B, T, nh, n_unmasked, y, att, att_T, kcache, vcache, Tmax);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int gpt_attention_decode(float* qkv32, int B, int nh, const int* pos_ptr, __nv_bfloat16* kcache,
                         __nv_bfloat16* vcache, int Tmax, __nv_bfloat16* y, float* att_rows, int Tatt, bool zero_consumed,
                         float* zero_buf, long long zero_count, cudaStream_t s, bool pdl, const LnFold* fold) {
  MGV_REQUIRE(zero_buf == nullptr || (B > 0 && zero_count % (static_cast<long long>(B) * nh) == 0),
              "attention: zero_count must be a multiple of the CTA count");
  MGV_REQUIRE(Tmax <= GPT_MAX_T, "attention: Tmax=%d exceeds %d", Tmax, GPT_MAX_T);
  if (B == 0) return MGV_OK;
  LaunchCfg lc(dim3(B * nh), dim3(AD_THREADS), 0, s, pdl);
  const int zero_per_cta = zero_buf ? static_cast<int>(zero_count / (static_cast<long long>(B) * nh)) : 0;
  const LnFold f = fold ? *fold : LnFold();
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, attn_decode_kernel, qkv32, nh, pos_ptr, kcache, vcache, Tmax, y, att_rows,
                                    Tatt, zero_consumed ? 1 : 0, zero_buf, zero_per_cta, f));
  return MGV_OK;
}

int gpt_gelu_bf16(float* h32, long long n, __nv_bfloat16* out, bool zero_consumed, cudaStream_t s, bool pdl) {
  MGV_REQUIRE(n % 4 == 0, "gelu: n");
  if (n == 0) return MGV_OK;
  const long long n4 = n / 4;
  static const int vec = getenv("MGV_GELU_VEC") ? atoi(getenv("MGV_GELU_VEC")) : GELU_VEC;   // threads cover `vec` float4 each
  const int threads = 256;   // 128 measured the same
  int blocks = static_cast<int>((n4 + static_cast<long long>(threads) * vec - 1) / (static_cast<long long>(threads) * vec));
  if (blocks < 1) blocks = 1;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  LaunchCfg lc(dim3(blocks), dim3(threads), 0, s, pdl);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, gelu_bf16_kernel, h32, n4, out, zero_consumed ? 1 : 0));
  return MGV_OK;
}

int gpt_sample_step(const SampleArgs& a, cudaStream_t s, bool pdl) {
  MGV_REQUIRE(a.done_counter && a.pos_ptr && a.logits_acc && a.seed_ptr, "sample: null state pointers");
  MGV_REQUIRE(a.V >= 1 && a.V <= SAMPLE_MAX_V, "sample: vocab=%d unsupported (<= %d)", a.V, SAMPLE_MAX_V);
  MGV_REQUIRE(a.C % 4 == 0, "sample: C=%d", a.C);
  MGV_REQUIRE(a.temperature > 0.f, "sample: temperature must be > 0");
  if (a.B == 0) return MGV_OK;
  const size_t smem = static_cast<size_t>(a.V) * 4;
  LaunchCfg lc(dim3(a.B), dim3(SAMPLE_THREADS), smem, s, pdl);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, sample_step_kernel, a));
  return MGV_OK;
}

}  // namespace mgv
