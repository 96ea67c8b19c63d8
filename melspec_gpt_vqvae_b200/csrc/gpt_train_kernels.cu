// Non-GEMM kernels of the minGPT training step (see gpt_train.cuh): dropout, attention forward / backward with
// log-sum-exp recomputation, LayerNorm / GELU / cross-entropy backward, transposes for the weight-gradient GEMMs,
// embedding backward and the fused AdamW update.
// reference: Lit_minGPT.training_step / shared_step transformer/minGPT.py:413-422, GPT.forward :168-199,
// CausalSelfAttention.forward :72-90, Block.forward :107-119, configure_optimizers :618-665.
#include "gpt_train.cuh"

namespace mgv {

namespace {

// ------------------------------------------------------------------ dropout mask
// Counter-based: keep(element) depends only on (seed, stream, element index), so the forward and the backward kernels
// (which visit the elements in different orders and groupings) regenerate identical masks and nothing is stored.
// lowbias32 (two multiply-xorshift rounds) twice; tests/ re-implement it in torch to build the same masks.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool drop_keep(const DropCfg& d, uint32_t stream, uint32_t idx) {
  uint32_t h = mix32(idx ^ d.seed_lo);
  h = mix32(h + stream * 0x9E3779B9u + d.seed_hi);
  return (h >> 8) >= d.thresh24;      // P(keep) = 1 - p
}

// ------------------------------------------------------------------ embedding (+ dropout)
__global__ void embed_train_kernel(const long long* __restrict__ idx, int T, int t, const long long* __restrict__ cls,
                                   const float* __restrict__ embedder, int m, const float* __restrict__ tok_emb,
                                   const float* __restrict__ pos_emb, int C, int vocab, int class_size,
                                   float* __restrict__ x_out, int* __restrict__ err_flag, DropCfg dc, float inv_keep) {
  const int row = blockIdx.x;   // b*T + p
  const int b = row / T, p = row - b * T;
  const float* src;
  if (p < m) {
    long long c = cls[b];
    if (c < 0 || c >= class_size) {
      atomicExch(err_flag, 1);
      c = 0;
    }
    src = embedder + c * C;
  } else {
    long long tok = idx[static_cast<long long>(b) * t + (p - m)];
    if (tok < 0 || tok >= vocab) {
      atomicExch(err_flag, 2);
      tok = 0;
    }
    src = tok_emb + tok * C;
  }
  const float* pe = pos_emb + static_cast<long long>(p) * C;
  float* o = x_out + static_cast<long long>(row) * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float v = src[i] + pe[i];
    if (dc.thresh24 != 0u) v = drop_keep(dc, DROP_STREAM_EMBD, static_cast<uint32_t>(row) * C + i) ? v * inv_keep : 0.f;
    o[i] = v;
  }
}

// x_out = resid + dropout(branch)     (resid_drop after proj, Dropout at the end of the mlp)
__global__ void resid_dropout_kernel(const float* __restrict__ resid, const float* __restrict__ branch, long long n,
                                     float* __restrict__ x_out, DropCfg dc, uint32_t stream, float inv_keep) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = drop_keep(dc, stream, static_cast<uint32_t>(i)) ? branch[i] * inv_keep : 0.f;
    x_out[i] = resid[i] + v;
  }
}

__global__ void gelu_fwd_kernel(const __nv_bfloat16* __restrict__ hpre, long long n, __nv_bfloat16* __restrict__ h) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    h[i] = __float2bfloat16(gelu_erf(__bfloat162float(hpre[i])));
}

// ------------------------------------------------------------------ gradient preparation
// One pass over a [R, N] gradient matrix that produces what the two GEMMs consuming it need:
//   g   [R, N]    bf16 row-major   (dgrad:  dX = g W)
//   gT  [N, Rpad] bf16             (wgrad:  dW = gT actT^T; contraction over the rows)
//   db  [N] += column sums         (bias gradient)
// and applies the element-wise backward of what produced the forward value:
//   MODE_DROP : src fp32, g = keep ? src / (1-p) : 0     (residual-branch dropout; p = 0: plain conversion)
//   MODE_GELU : src bf16 (dh), aux = hpre bf16, g = src * gelu'(aux)
//   MODE_COPY : src bf16, g = src (no g output needed: only gT and db)
// 32 x 32 tiles through shared memory; a CTA walks TILE_ROWS row tiles of one 32-column strip.
enum { GP_DROP = 0, GP_GELU = 1, GP_COPY = 2 };
constexpr int GP_ROW_TILES = 8;   // 256 rows per CTA

__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <int MODE>
__global__ void __launch_bounds__(256)
grad_prep_kernel(const void* __restrict__ src, const __nv_bfloat16* __restrict__ aux, int R, int N, int Rpad,
                 __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ gT, float* __restrict__ db, DropCfg dc,
                 uint32_t stream, float inv_keep) {
  __shared__ float tile[32][33];
  __shared__ float colsum[8][32];
  const int c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  float csum = 0.f;
  for (int rt = 0; rt < GP_ROW_TILES; ++rt) {
    const int r0 = (blockIdx.y * GP_ROW_TILES + rt) * 32;
    if (r0 >= R) break;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + 8 * k, c = c0 + tx;
      float v = 0.f;
      if (r < R && c < N) {
        const long long i = static_cast<long long>(r) * N + c;
        if (MODE == GP_DROP) {
          v = static_cast<const float*>(src)[i];
          if (dc.thresh24 != 0u) v = drop_keep(dc, stream, static_cast<uint32_t>(i)) ? v * inv_keep : 0.f;
        } else if (MODE == GP_GELU) {
          v = __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]) * gelu_grad(__bfloat162float(aux[i]));
        } else {
          v = __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]);
        }
        const __nv_bfloat16 vb = __float2bfloat16(v);
        if (MODE != GP_COPY) g[i] = vb;
        v = __bfloat162float(vb);          // the GEMMs see the rounded value; keep the bias gradient consistent with it
      }
      tile[ty + 8 * k][tx] = v;
      csum += v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k, r = r0 + tx;
      if (c < N && r < R) gT[static_cast<long long>(c) * Rpad + r] = __float2bfloat16(tile[tx][ty + 8 * k]);
    }
    __syncthreads();
  }
  if (db != nullptr) {
    colsum[ty][tx] = csum;
    __syncthreads();
    if (ty == 0 && c0 + tx < N) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += colsum[k][tx];
      atomicAdd(db + c0 + tx, s);
    }
  }
}

// plain bf16 transpose [R, N] -> [N, Rpad] (forward activations for the weight-gradient GEMMs)
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int R, int N, long long ld_src, int Rpad,
                      __nv_bfloat16* __restrict__ dst) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < R && c < N) ? src[static_cast<long long>(r) * ld_src + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (c < N && r < R) dst[static_cast<long long>(c) * Rpad + r] = tile[tx][ty + 8 * k];
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// dx_io[row] += rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;   dgamma += sum_rows dy * xhat;
// dbeta += sum_rows dy.  The statistics are recomputed from the saved input row.  A CTA = 8 warps x LNB_ROWS rows each;
// per-lane partial dgamma / dbeta live in registers, are folded across the warps in shared memory and leave the CTA
// as one atomicAdd per column.

template <int LNB_MAX_V4>        // float4 chunks per lane: C <= 128 * LNB_MAX_V4
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma, int rows,
                     int C, float* __restrict__ dx_io, int accumulate, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int LNB_ROWS) {
  extern __shared__ __align__(16) float lnb_smem[];   // [8 warps][2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = C / 4;
  float4 dg[LNB_MAX_V4], dbt[LNB_MAX_V4];
#pragma unroll
  for (int j = 0; j < LNB_MAX_V4; ++j) dg[j] = dbt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  for (int rr = 0; rr < LNB_ROWS; ++rr) {
    const int row = (blockIdx.x * 8 + warp) * LNB_ROWS + rr;
    if (row >= rows) break;
    const float4* x4 = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
    const float4* d4 = reinterpret_cast<const float4*>(dy + static_cast<long long>(row) * C);
    float4 xv[LNB_MAX_V4], dv[LNB_MAX_V4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LNB_MAX_V4; ++j) {
      const int i = lane + 32 * j;
      if (i < nv) {
        xv[j] = x4[i];
        dv[j] = d4[i];
        s += (xv[j].x + xv[j].y) + (xv[j].z + xv[j].w);
      }
    }
    const float mean = warp_sum(s) / static_cast<float>(C);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < LNB_MAX_V4; ++j) {
      const int i = lane + 32 * j;
      if (i < nv) {
        const float a = xv[j].x - mean, b = xv[j].y - mean, c = xv[j].z - mean, d = xv[j].w - mean;
        ss += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(C) + 1e-5f);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < LNB_MAX_V4; ++j) {
      const int i = lane + 32 * j;
      if (i < nv) {
        const float4 gm = __ldg(g4 + i);
        // xv <- xhat, dv stays dy; accumulate parameter gradients
        xv[j].x = (xv[j].x - mean) * rstd; xv[j].y = (xv[j].y - mean) * rstd;
        xv[j].z = (xv[j].z - mean) * rstd; xv[j].w = (xv[j].w - mean) * rstd;
        dg[j].x += dv[j].x * xv[j].x; dg[j].y += dv[j].y * xv[j].y; dg[j].z += dv[j].z * xv[j].z; dg[j].w += dv[j].w * xv[j].w;
        dbt[j].x += dv[j].x; dbt[j].y += dv[j].y; dbt[j].z += dv[j].z; dbt[j].w += dv[j].w;
        dv[j].x *= gm.x; dv[j].y *= gm.y; dv[j].z *= gm.z; dv[j].w *= gm.w;      // dv <- g = dy * gamma
        sg += (dv[j].x + dv[j].y) + (dv[j].z + dv[j].w);
        sgx += (dv[j].x * xv[j].x + dv[j].y * xv[j].y) + (dv[j].z * xv[j].z + dv[j].w * xv[j].w);
      }
    }
    const float mg = warp_sum(sg) / static_cast<float>(C), mgx = warp_sum(sgx) / static_cast<float>(C);
    float4* o4 = reinterpret_cast<float4*>(dx_io + static_cast<long long>(row) * C);
#pragma unroll
    for (int j = 0; j < LNB_MAX_V4; ++j) {
      const int i = lane + 32 * j;
      if (i < nv) {
        float4 r;
        r.x = rstd * (dv[j].x - mg - xv[j].x * mgx);
        r.y = rstd * (dv[j].y - mg - xv[j].y * mgx);
        r.z = rstd * (dv[j].z - mg - xv[j].z * mgx);
        r.w = rstd * (dv[j].w - mg - xv[j].w * mgx);
        if (accumulate) {
          const float4 old = o4[i];
          r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
        }
        o4[i] = r;
      }
    }
  }
  // fold the warps' parameter-gradient partials: every warp parks its row of 2C partial sums in shared memory, then
  // thread t adds the 8 warps' values of its columns and issues ONE global atomic per column and CTA
  float* mine = lnb_smem + static_cast<size_t>(warp) * 2 * C;
#pragma unroll
  for (int j = 0; j < LNB_MAX_V4; ++j) {
    const int i = lane + 32 * j;
    if (i < nv) {
      *reinterpret_cast<float4*>(mine + 4 * i) = dg[j];
      *reinterpret_cast<float4*>(mine + C + 4 * i) = dbt[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += lnb_smem[static_cast<size_t>(w) * 2 * C + i];
    atomicAdd((i < C ? dgamma : dbeta - C) + i, s);
  }
}

// ------------------------------------------------------------------ cross entropy forward + backward
// loss_sum += sum_rows (logsumexp - logit[target]) * inv_rows;   dlogits = (softmax - onehot) * inv_rows   (bf16)
// (F.cross_entropy with mean reduction, minGPT.py:416).  One warp per row.
__global__ void __launch_bounds__(128)
ce_fwd_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ targets, long long rows, int V,
                  float inv_rows, float* __restrict__ loss_sum, __nv_bfloat16* __restrict__ dlogits,
                  int* __restrict__ err_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (r >= rows) return;
  const float* row = logits + r * V;
  float mx = -INFINITY;
  for (int i = lane; i < V; i += 32) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < V; i += 32) sum += expf(row[i] - mx);
  sum = warp_sum(sum);
  long long t = targets[r];
  if (t < 0 || t >= V) {
    if (lane == 0) *err_flag = 2;
    t = 0;
  }
  const float inv = 1.0f / sum;
  for (int i = lane; i < V; i += 32) {
    const float p = expf(row[i] - mx) * inv;
    dlogits[r * V + i] = __float2bfloat16((p - (i == t ? 1.f : 0.f)) * inv_rows);
  }
  if (lane == 0) atomicAdd(loss_sum, ((mx + logf(sum)) - row[t]) * inv_rows);
}

// ------------------------------------------------------------------ embedding backward
// dx [B*T, C] is the gradient of the (dropped-out) sum tok_emb[idx] + pos_emb[p]: scatter-add into the three tables.
__global__ void embed_bwd_kernel(const float* __restrict__ dx, const long long* __restrict__ idx, int T, int t,
                                 const long long* __restrict__ cls, int m, int C, float* __restrict__ d_tok,
                                 float* __restrict__ d_pos, float* __restrict__ d_embedder, DropCfg dc, float inv_keep) {
  const int row = blockIdx.x;
  const int b = row / T, p = row - b * T;
  float* dst = (p < m) ? d_embedder + cls[b] * C : d_tok + idx[static_cast<long long>(b) * t + (p - m)] * C;
  float* dpe = d_pos + static_cast<long long>(p) * C;
  const float* src = dx + static_cast<long long>(row) * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float v = src[i];
    if (dc.thresh24 != 0u) v = drop_keep(dc, DROP_STREAM_EMBD, static_cast<uint32_t>(row) * C + i) ? v * inv_keep : 0.f;
    atomicAdd(dst + i, v);
    atomicAdd(dpe + i, v);
  }
}

// ------------------------------------------------------------------ attention (training): forward with log-sum-exp
// Same structure as attn_prefill_kernel (gpt_kernels.cu): one CTA per (sequence, head), K and V staged once in shared
// memory (rows padded to 144 B), mma.sync m16n8k16 bf16 -> fp32, 9 warps walk the 16-row query blocks in
// causal-balanced pairs, single pass with an online softmax.  Additions: attention dropout (minGPT.py:84, mask from the
// counter hash: the row sum uses the UNdropped probabilities, the value accumulation the dropped ones) and the per-row
// log2-sum-exp, from which the backward kernels recompute the probabilities (nothing of size T x T is stored).
constexpr int TA_BN = 64;
constexpr int TA_LD = 72;
constexpr int TA_WARPS = 9;
constexpr int TA_THREADS = TA_WARPS * 32;
constexpr int TA_MAXK = 320;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r0), "=r"(r1)
               : "r"(smem_u32(smem_row)));
}

// stage `rows` x 64 bf16 (row stride ld elements in global memory) into shared memory rows of TA_LD, zero beyond `rows`
__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int rows, int kpad) {
  for (int i = threadIdx.x; i < kpad * 8; i += TA_THREADS) {
    const int r = i >> 3, part = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows) v = *reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * ld + part * 8);
    *reinterpret_cast<uint4*>(dst + r * TA_LD + part * 8) = v;
  }
}
// A-operand fragments (16 rows x 64 k) straight from global memory; rows beyond `rows` mirror the last row
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const __nv_bfloat16* src, long long ld, int rbase,
                                             int rows, int g, int tq) {
  const int r0 = min(rbase + g, rows - 1), r1 = min(rbase + g + 8, rows - 1);
  const __nv_bfloat16* p0 = src + static_cast<long long>(r0) * ld + tq * 2;
  const __nv_bfloat16* p1 = src + static_cast<long long>(r1) * ld + tq * 2;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = *reinterpret_cast<const uint32_t*>(p0 + kk * 16);
    a[kk][1] = *reinterpret_cast<const uint32_t*>(p1 + kk * 16);
    a[kk][2] = *reinterpret_cast<const uint32_t*>(p0 + kk * 16 + 8);
    a[kk][3] = *reinterpret_cast<const uint32_t*>(p1 + kk * 16 + 8);
  }
}
// acc[nb][e] = sum_k A[16 x 64] * Bs[(jb*64 + nb*8 + n)][k]   (Bs rows in shared memory, 64 of them)
__device__ __forceinline__ void mma_block(float (&acc)[8][4], const uint32_t (&a)[4][4], const __nv_bfloat16* Bs, int jb,
                                          int g, int tq) {
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
    const __nv_bfloat16* kb = Bs + (jb * TA_BN + nb * 8 + g) * TA_LD + tq * 2;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kb + kk * 16);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kb + kk * 16 + 8);
      mma16816(acc[nb], a[kk], b0, b1);
    }
  }
}
// o[nd][e] += P[16 x 64 (block jb)] * Bs[(jb*64 + k)][nd*8 + n]   (P from accumulator fragments, Bs rows = contraction index)
__device__ __forceinline__ void mma_pv(float (&o)[8][4], const float (&p)[8][4], const __nv_bfloat16* Bs, int jb, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t ap[4];
    ap[0] = pack_bf16x2(p[2 * ks][0], p[2 * ks][1]);
    ap[1] = pack_bf16x2(p[2 * ks][2], p[2 * ks][3]);
    ap[2] = pack_bf16x2(p[2 * ks + 1][0], p[2 * ks + 1][1]);
    ap[3] = pack_bf16x2(p[2 * ks + 1][2], p[2 * ks + 1][3]);
    const __nv_bfloat16* vrow = Bs + (jb * TA_BN + ks * 16 + (lane & 15)) * TA_LD;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      uint32_t b0, b1;
      ldsm_x2_trans(b0, b1, vrow + nd * 8);
      mma16816(o[nd], ap, b0, b1);
    }
  }
}

__device__ __forceinline__ bool att_allowed(int row, int key, int T, int n_unmasked) {
  // mask[i][j] = tril, plus the unmasked prefix block (minGPT.py:65-68, :82)
  return key < T && row < T && (key <= row || (row < n_unmasked && key < n_unmasked));
}
__device__ __forceinline__ uint32_t att_drop_index(int bh, int T, int row, int key) {
  return (static_cast<uint32_t>(bh) * T + row) * T + key;
}

__global__ void __launch_bounds__(TA_THREADS, 2)
attn_train_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int T, int nh, int n_unmasked, __nv_bfloat16* __restrict__ y,
                      float* __restrict__ lse2, DropCfg dc, uint32_t stream, float inv_keep) {
  extern __shared__ __align__(16) __nv_bfloat16 ta_smem[];
  const int kpad = ((T + TA_BN - 1) / TA_BN) * TA_BN;
  __nv_bfloat16* sK = ta_smem;
  __nv_bfloat16* sV = sK + kpad * TA_LD;
  const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
  const int C = nh * GPT_HEAD_DIM;
  const long long ld = 3 * C;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ld + h * GPT_HEAD_DIM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  stage_rows(sK, base + C, ld, T, kpad);
  stage_rows(sV, base + 2 * C, ld, T, kpad);
  __syncthreads();
  const int nrb = (T + 15) / 16;
  const float scale2 = 1.4426950408889634f / sqrtf(static_cast<float>(GPT_HEAD_DIM));
  for (int turn = 0; turn < 2; ++turn) {
    const int rb = (turn == 0) ? warp : nrb - 1 - warp;
    if (turn == 0 ? (2 * warp >= nrb) : (rb <= warp)) continue;
    const int rbase = rb * 16;
    uint32_t aq[4][4];
    load_a_frags(aq, base, ld, rbase, T, g, tq);
    const int row0 = rbase + g, row1 = rbase + g + 8;
    const int wlast = min(rbase + 15, T - 1);
    const int wkmax = (rbase < n_unmasked) ? max(wlast + 1, min(n_unmasked, T)) : wlast + 1;
    const int wnkb = (wkmax + TA_BN - 1) / TA_BN;
    float o[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int jb = 0; jb < wnkb; ++jb) {
      float sc[8][4];
      mma_block(sc, aq, sK, jb, g, tq);
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = jb * TA_BN + nb * 8 + tq * 2 + (e & 1);
          const int row = (e < 2) ? row0 : row1;
          // padded rows (>= T) mirror the last row's mask so that their maxima stay finite; they are never stored
          sc[nb][e] = att_allowed(min(row, T - 1), key, T, n_unmasked) ? sc[nb][e] * scale2 : -INFINITY;
          if (e < 2) bm0 = fmaxf(bm0, sc[nb][e]); else bm1 = fmaxf(bm1, sc[nb][e]);
        }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
      const float a0 = exp2f(m0 - n0), a1 = exp2f(m1 - n1);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pr = exp2f(sc[nb][e] - ((e < 2) ? n0 : n1));
          if (e < 2) s0 += pr; else s1 += pr;
          float pd = pr;
          if (dc.thresh24 != 0u) {
            const int key = jb * TA_BN + nb * 8 + tq * 2 + (e & 1);
            const int row = (e < 2) ? row0 : row1;
            pd = drop_keep(dc, stream, att_drop_index(bh, T, min(row, T - 1), key)) ? pr * inv_keep : 0.f;
          }
          sc[nb][e] = pd;
        }
      l0 = l0 * a0 + s0;
      l1 = l1 * a1 + s1;
      m0 = n0;
      m1 = n1;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        o[nd][0] *= a0; o[nd][1] *= a0; o[nd][2] *= a1; o[nd][3] *= a1;
      }
      mma_pv(o, sc, sV, jb, lane);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    if (tq == 0) {
      if (row0 < T) lse2[static_cast<long long>(bh) * T + row0] = m0 + log2f(l0);
      if (row1 < T) lse2[static_cast<long long>(bh) * T + row1] = m1 + log2f(l1);
    }
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      const int dim = nd * 8 + tq * 2;
      if (row0 < T)
        *reinterpret_cast<uint32_t*>(y + (static_cast<long long>(b) * T + row0) * C + h * GPT_HEAD_DIM + dim) =
            pack_bf16x2(o[nd][0] * inv0, o[nd][1] * inv0);
      if (row1 < T)
        *reinterpret_cast<uint32_t*>(y + (static_cast<long long>(b) * T + row1) * C + h * GPT_HEAD_DIM + dim) =
            pack_bf16x2(o[nd][2] * inv1, o[nd][3] * inv1);
    }
  }
}

// ------------------------------------------------------------------ attention backward, part 1: dQ (and delta)
// Warps walk the query blocks.  For every key block: S = Q K^T (recomputed), P = exp2(S - lse), dPd = dY V^T,
// dP = dropout-backward(dPd), dS = P * (dP - delta) with delta = rowsum(dY * Y), dQ += dS K.  K and V rows in shared
// memory; Q / dY / Y fragments from global memory.  delta is written for part 2.
__global__ void __launch_bounds__(TA_THREADS, 1)   // Q, dY fragments + three 16 x 64 fp32 tiles per thread: one CTA per SM
attn_train_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ yv,
                         const __nv_bfloat16* __restrict__ dy, const float* __restrict__ lse2, int T, int nh,
                         int n_unmasked, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ delta, DropCfg dc,
                         uint32_t stream, float inv_keep) {
  extern __shared__ __align__(16) __nv_bfloat16 ta_smem[];
  const int kpad = ((T + TA_BN - 1) / TA_BN) * TA_BN;
  __nv_bfloat16* sK = ta_smem;
  __nv_bfloat16* sV = sK + kpad * TA_LD;
  const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
  const int C = nh * GPT_HEAD_DIM;
  const long long ld = 3 * C;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ld + h * GPT_HEAD_DIM;
  const __nv_bfloat16* ybase = yv + static_cast<long long>(b) * T * C + h * GPT_HEAD_DIM;
  const __nv_bfloat16* dybase = dy + static_cast<long long>(b) * T * C + h * GPT_HEAD_DIM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  stage_rows(sK, base + C, ld, T, kpad);
  stage_rows(sV, base + 2 * C, ld, T, kpad);
  __syncthreads();
  const int nrb = (T + 15) / 16;
  const float scale2 = 1.4426950408889634f / sqrtf(static_cast<float>(GPT_HEAD_DIM));
  const float scale = 1.0f / sqrtf(static_cast<float>(GPT_HEAD_DIM));
  for (int turn = 0; turn < 2; ++turn) {
    const int rb = (turn == 0) ? warp : nrb - 1 - warp;
    if (turn == 0 ? (2 * warp >= nrb) : (rb <= warp)) continue;
    const int rbase = rb * 16;
    uint32_t aq[4][4], ady[4][4], ayv[4][4];
    load_a_frags(aq, base, ld, rbase, T, g, tq);
    load_a_frags(ady, dybase, C, rbase, T, g, tq);
    load_a_frags(ayv, ybase, C, rbase, T, g, tq);
    const int row0 = rbase + g, row1 = rbase + g + 8;
    // delta = rowsum(dY * Y): this thread holds 16 of the 64 dims of rows g (regs 0, 2) and g+8 (regs 1, 3)
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float2 a = unpack_bf16x2(ady[kk][r]), c = unpack_bf16x2(ayv[kk][r]);
        const float t = a.x * c.x + a.y * c.y;
        if (r & 1) d1 += t; else d0 += t;
      }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    if (tq == 0) {
      if (row0 < T) delta[static_cast<long long>(bh) * T + row0] = d0;
      if (row1 < T) delta[static_cast<long long>(bh) * T + row1] = d1;
    }
    const float ls0 = lse2[static_cast<long long>(bh) * T + min(row0, T - 1)];
    const float ls1 = lse2[static_cast<long long>(bh) * T + min(row1, T - 1)];
    const int wlast = min(rbase + 15, T - 1);
    const int wkmax = (rbase < n_unmasked) ? max(wlast + 1, min(n_unmasked, T)) : wlast + 1;
    const int wnkb = (wkmax + TA_BN - 1) / TA_BN;
    float dq[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) dq[nd][0] = dq[nd][1] = dq[nd][2] = dq[nd][3] = 0.f;
    for (int jb = 0; jb < wnkb; ++jb) {
      float sc[8][4], dp[8][4];
      mma_block(sc, aq, sK, jb, g, tq);
      mma_block(dp, ady, sV, jb, g, tq);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = jb * TA_BN + nb * 8 + tq * 2 + (e & 1);
          const int row = (e < 2) ? row0 : row1;
          float ds = 0.f;
          if (att_allowed(row, key, T, n_unmasked)) {
            const float pr = exp2f(sc[nb][e] * scale2 - ((e < 2) ? ls0 : ls1));
            float dpr = dp[nb][e];
            if (dc.thresh24 != 0u) dpr = drop_keep(dc, stream, att_drop_index(bh, T, row, key)) ? dpr * inv_keep : 0.f;
            ds = pr * (dpr - ((e < 2) ? d0 : d1));
          }
          sc[nb][e] = ds;
        }
      mma_pv(dq, sc, sK, jb, lane);
    }
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      const int dim = nd * 8 + tq * 2;
      if (row0 < T)
        *reinterpret_cast<uint32_t*>(dqkv + (static_cast<long long>(b) * T + row0) * ld + h * GPT_HEAD_DIM + dim) =
            pack_bf16x2(dq[nd][0] * scale, dq[nd][1] * scale);
      if (row1 < T)
        *reinterpret_cast<uint32_t*>(dqkv + (static_cast<long long>(b) * T + row1) * ld + h * GPT_HEAD_DIM + dim) =
            pack_bf16x2(dq[nd][2] * scale, dq[nd][3] * scale);
    }
  }
}

// ------------------------------------------------------------------ attention backward, part 2: dK, dV
// The transposed problem: warps walk the KEY blocks; for every block of 64 query rows S^T = K Q^T, P^T, dPd^T = V dY^T,
// dS^T as above (lse / delta now index the columns), dK += dS^T Q, dV += Pd^T dY.  Q and dY rows in shared memory.
__global__ void __launch_bounds__(TA_THREADS, 1)
attn_train_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dy,
                          const float* __restrict__ lse2, const float* __restrict__ delta, int T, int nh, int n_unmasked,
                          __nv_bfloat16* __restrict__ dqkv, DropCfg dc, uint32_t stream, float inv_keep) {
  extern __shared__ __align__(16) __nv_bfloat16 ta_smem[];
  const int kpad = ((T + TA_BN - 1) / TA_BN) * TA_BN;
  __nv_bfloat16* sQ = ta_smem;
  __nv_bfloat16* sdY = sQ + kpad * TA_LD;
  float* s_lse = reinterpret_cast<float*>(sdY + kpad * TA_LD);   // [kpad]
  float* s_del = s_lse + kpad;                                   // [kpad]
  const int bh = blockIdx.x, b = bh / nh, h = bh - b * nh;
  const int C = nh * GPT_HEAD_DIM;
  const long long ld = 3 * C;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ld + h * GPT_HEAD_DIM;
  const __nv_bfloat16* dybase = dy + static_cast<long long>(b) * T * C + h * GPT_HEAD_DIM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  stage_rows(sQ, base, ld, T, kpad);
  stage_rows(sdY, dybase, C, T, kpad);
  for (int i = threadIdx.x; i < kpad; i += TA_THREADS) {
    s_lse[i] = (i < T) ? lse2[static_cast<long long>(bh) * T + i] : 0.f;
    s_del[i] = (i < T) ? delta[static_cast<long long>(bh) * T + i] : 0.f;
  }
  __syncthreads();
  const int nrb = (T + 15) / 16;
  const int nqb = kpad / TA_BN;
  const float scale2 = 1.4426950408889634f / sqrtf(static_cast<float>(GPT_HEAD_DIM));
  const float scale = 1.0f / sqrtf(static_cast<float>(GPT_HEAD_DIM));
  for (int turn = 0; turn < 2; ++turn) {
    // key block w sees the query blocks from its own onwards: pair w with nrb-1-w like the forward
    const int kbk = (turn == 0) ? warp : nrb - 1 - warp;
    if (turn == 0 ? (2 * warp >= nrb) : (kbk <= warp)) continue;
    const int kbase = kbk * 16;
    uint32_t ak[4][4], av[4][4];
    load_a_frags(ak, base + C, ld, kbase, T, g, tq);
    load_a_frags(av, base + 2 * C, ld, kbase, T, g, tq);
    const int key0 = kbase + g, key1 = kbase + g + 8;
    // first 64-row query block that can see these keys (all of them when the keys lie in the unmasked prefix)
    const int ib0 = (kbase < n_unmasked) ? 0 : kbase / TA_BN;
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      dk[nd][0] = dk[nd][1] = dk[nd][2] = dk[nd][3] = 0.f;
      dv[nd][0] = dv[nd][1] = dv[nd][2] = dv[nd][3] = 0.f;
    }
    for (int ib = ib0; ib < nqb; ++ib) {
      float st[8][4], dpt[8][4];
      mma_block(st, ak, sQ, ib, g, tq);       // S^T[key, row]
      mma_block(dpt, av, sdY, ib, g, tq);     // dPd^T[key, row]
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = ib * TA_BN + nb * 8 + tq * 2 + (e & 1);
          const int key = (e < 2) ? key0 : key1;
          float ds = 0.f, pd = 0.f;
          if (att_allowed(row, key, T, n_unmasked)) {
            const float pr = exp2f(st[nb][e] * scale2 - s_lse[row]);
            float dpr = dpt[nb][e];
            pd = pr;
            if (dc.thresh24 != 0u) {
              const bool keep = drop_keep(dc, stream, att_drop_index(bh, T, row, key));
              dpr = keep ? dpr * inv_keep : 0.f;
              pd = keep ? pr * inv_keep : 0.f;
            }
            ds = pr * (dpr - s_del[row]);
          }
          st[nb][e] = ds;
          dpt[nb][e] = pd;
        }
      mma_pv(dk, st, sQ, ib, lane);
      mma_pv(dv, dpt, sdY, ib, lane);
    }
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      const int dim = nd * 8 + tq * 2;
      if (key0 < T) {
        __nv_bfloat16* o = dqkv + (static_cast<long long>(b) * T + key0) * ld + h * GPT_HEAD_DIM + dim;
        *reinterpret_cast<uint32_t*>(o + C) = pack_bf16x2(dk[nd][0] * scale, dk[nd][1] * scale);
        *reinterpret_cast<uint32_t*>(o + 2 * C) = pack_bf16x2(dv[nd][0], dv[nd][1]);
      }
      if (key1 < T) {
        __nv_bfloat16* o = dqkv + (static_cast<long long>(b) * T + key1) * ld + h * GPT_HEAD_DIM + dim;
        *reinterpret_cast<uint32_t*>(o + C) = pack_bf16x2(dk[nd][2] * scale, dk[nd][3] * scale);
        *reinterpret_cast<uint32_t*>(o + 2 * C) = pack_bf16x2(dv[nd][2], dv[nd][3]);
      }
    }
  }
}

// ------------------------------------------------------------------ fused AdamW (torch.optim.AdamW semantics)
// One pass over the flat fp32 parameter buffer: p *= 1 - lr*wd (decay segments only), Adam moments, bias-corrected
// update, and the refreshed inference copies (fp32 for LayerNorm / bias / embedding tensors, bf16 for Linear weights).
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             const AdamSeg* __restrict__ segs, const int2* __restrict__ chunks, float lr, float beta1, float beta2,
             float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  const int2 ck = chunks[blockIdx.x];
  const AdamSeg sg = segs[ck.x];
  const long long lo = static_cast<long long>(ck.y) * ADAM_CHUNK;
  const long long hi = (lo + ADAM_CHUNK < sg.numel) ? lo + ADAM_CHUNK : sg.numel;
  const float decay = sg.decay ? 1.0f - lr * wd : 1.0f;
  const float step = lr / bc1;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const long long o = sg.offset + i;
    const float gr = g[o] * grad_scale;
    float pv = p[o] * decay;
    const float mm = beta1 * m[o] + (1.0f - beta1) * gr;
    const float vv = beta2 * v[o] + (1.0f - beta2) * gr * gr;
    m[o] = mm;
    v[o] = vv;
    pv -= step * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    p[o] = pv;
    if (sg.dst_f32) sg.dst_f32[i] = pv;
    if (sg.dst_bf16) sg.dst_bf16[i] = __float2bfloat16(pv);
  }
}

__global__ void drop_mask_kernel(DropCfg dc, uint32_t stream, long long n, unsigned char* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = drop_keep(dc, stream, static_cast<uint32_t>(i)) ? 1 : 0;
}

int grid_for(long long n, int threads, int per_sm = 8) {
  long long b = (n + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

DropCfg make_drop(float p, unsigned long long seed) {
  DropCfg d;
  d.seed_lo = static_cast<uint32_t>(seed);
  d.seed_hi = static_cast<uint32_t>(seed >> 32);
  d.thresh24 = 0u;
  if (p > 0.f) {
    double t = static_cast<double>(p) * 16777216.0;
    if (t > 16777215.0) t = 16777215.0;
    d.thresh24 = static_cast<uint32_t>(t + 0.5);
    if (d.thresh24 == 0u) d.thresh24 = 1u;
  }
  return d;
}
// exact keep probability of the integer threshold (so that E[mask / keep] == 1)
float drop_inv_keep(const DropCfg& d) { return 16777216.0f / static_cast<float>(16777216u - d.thresh24); }

int train_drop_mask(const DropCfg& dc, unsigned stream_id, long long n, unsigned char* out, cudaStream_t s) {
  drop_mask_kernel<<<grid_for(n, 256), 256, 0, s>>>(dc, stream_id, n, out);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_embed(const long long* idx, int B, int T, int t, const long long* cls, const float* embedder, int m,
                const float* tok_emb, const float* pos_emb, int C, int vocab, int class_size, float* x_out, int* err_flag,
                const DropCfg& dc, cudaStream_t s) {
  MGV_REQUIRE(m == 0 || (cls && embedder), "train_embed: prefix without class embedder");
  embed_train_kernel<<<B * T, 256, 0, s>>>(idx, T, t, cls, embedder, m, tok_emb, pos_emb, C, vocab, class_size, x_out,
                                            err_flag, dc, drop_inv_keep(dc));
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_resid_dropout(const float* resid, const float* branch, long long n, float* x_out, const DropCfg& dc,
                        unsigned stream_id, cudaStream_t s) {
  resid_dropout_kernel<<<grid_for(n, 256), 256, 0, s>>>(resid, branch, n, x_out, dc, stream_id, drop_inv_keep(dc));
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_gelu_fwd(const __nv_bfloat16* hpre, long long n, __nv_bfloat16* h, cudaStream_t s) {
  gelu_fwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(hpre, n, h);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_grad_prep(int mode, const void* src, const __nv_bfloat16* aux, int R, int N, int Rpad, __nv_bfloat16* g,
                    __nv_bfloat16* gT, float* db, const DropCfg& dc, unsigned stream_id, cudaStream_t s) {
  MGV_REQUIRE(src && gT && R >= 1 && N >= 1 && Rpad >= R, "grad_prep: bad arguments");
  dim3 grid(ceil_div(N, 32), ceil_div(R, 32 * GP_ROW_TILES));
  const float ik = drop_inv_keep(dc);
  if (mode == GP_DROP) {
    MGV_REQUIRE(g, "grad_prep: dropout mode needs the row-major output");
    grad_prep_kernel<GP_DROP><<<grid, 256, 0, s>>>(src, aux, R, N, Rpad, g, gT, db, dc, stream_id, ik);
  } else if (mode == GP_GELU) {
    MGV_REQUIRE(g && aux, "grad_prep: gelu mode needs hpre and the row-major output");
    grad_prep_kernel<GP_GELU><<<grid, 256, 0, s>>>(src, aux, R, N, Rpad, g, gT, db, dc, stream_id, ik);
  } else {
    grad_prep_kernel<GP_COPY><<<grid, 256, 0, s>>>(src, aux, R, N, Rpad, g, gT, db, dc, stream_id, ik);
  }
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_transpose(const __nv_bfloat16* src, int R, int N, long long ld_src, int Rpad, __nv_bfloat16* dst, cudaStream_t s) {
  dim3 grid(ceil_div(N, 32), ceil_div(R, 32));
  transpose_bf16_kernel<<<grid, 256, 0, s>>>(src, R, N, ld_src, Rpad, dst);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_layernorm_bwd(const float* dy, const float* x, const float* gamma, int rows, int C, float* dx_io, bool accumulate,
                        float* dgamma, float* dbeta, cudaStream_t s) {
  MGV_REQUIRE(C % 4 == 0 && C <= 16 * 128, "layernorm backward: C=%d unsupported", C);
  // rows per warp: enough CTAs to fill the GPU twice over at small batches, fewer column atomics at large ones
  int rpw = rows / (8 * 2 * num_sms());
  if (rpw < 1) rpw = 1;
  if (rpw > 8) rpw = 8;
  const int grid = ceil_div(rows, 8 * rpw);
  const size_t smem = static_cast<size_t>(8) * 2 * C * sizeof(float);
  const int acc = accumulate ? 1 : 0;
  static unsigned long long attr_mask = 0;
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4));
  }
  if (C <= 256) layernorm_bwd_kernel<2><<<grid, 256, smem, s>>>(dy, x, gamma, rows, C, dx_io, acc, dgamma, dbeta, rpw);
  else if (C <= 1024) layernorm_bwd_kernel<8><<<grid, 256, smem, s>>>(dy, x, gamma, rows, C, dx_io, acc, dgamma, dbeta, rpw);
  else layernorm_bwd_kernel<16><<<grid, 256, smem, s>>>(dy, x, gamma, rows, C, dx_io, acc, dgamma, dbeta, rpw);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_ce(const float* logits, const long long* targets, long long rows, int V, float* loss_sum,
             __nv_bfloat16* dlogits, int* err_flag, cudaStream_t s) {
  const long long blocks = (rows + 3) / 4;
  ce_fwd_bwd_kernel<<<static_cast<unsigned>(blocks), 128, 0, s>>>(logits, targets, rows, V, 1.0f / static_cast<float>(rows),
                                                                   loss_sum, dlogits, err_flag);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_embed_bwd(const float* dx, const long long* idx, int B, int T, int t, const long long* cls, int m, int C,
                    float* d_tok, float* d_pos, float* d_embedder, const DropCfg& dc, cudaStream_t s) {
  embed_bwd_kernel<<<B * T, 256, 0, s>>>(dx, idx, T, t, cls, m, C, d_tok, d_pos, d_embedder, dc, drop_inv_keep(dc));
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

namespace {
int set_attn_attrs() {
  static unsigned long long attr_mask = 0;
  if (first_use_on_this_device(attr_mask)) {
    const int big = (2 * TA_MAXK * TA_LD) * 2 + 2 * TA_MAXK * 4;
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_train_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_train_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_train_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_train_bwd_dq_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_train_bwd_dkv_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  return MGV_OK;
}
}  // namespace

int train_attn_fwd(const __nv_bfloat16* qkv, int B, int T, int nh, int n_unmasked, __nv_bfloat16* y, float* lse2,
                   const DropCfg& dc, unsigned stream_id, cudaStream_t s) {
  MGV_REQUIRE(T >= 1 && T <= GPT_MAX_T && (T + 15) / 16 <= 2 * TA_WARPS, "train attention: T=%d exceeds %d", T, GPT_MAX_T);
  MGV_TRY(set_attn_attrs());
  const int kpad = ceil_div(T, TA_BN) * TA_BN;
  const size_t smem = static_cast<size_t>(2 * kpad) * TA_LD * 2;
  attn_train_fwd_kernel<<<B * nh, TA_THREADS, smem, s>>>(qkv, T, nh, n_unmasked, y, lse2, dc, stream_id, drop_inv_keep(dc));
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_attn_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* y, const __nv_bfloat16* dy, const float* lse2, int B, int T,
                   int nh, int n_unmasked, __nv_bfloat16* dqkv, float* delta, const DropCfg& dc, unsigned stream_id,
                   cudaStream_t s) {
  MGV_REQUIRE(T >= 1 && T <= GPT_MAX_T && (T + 15) / 16 <= 2 * TA_WARPS, "train attention: T=%d exceeds %d", T, GPT_MAX_T);
  MGV_TRY(set_attn_attrs());
  const int kpad = ceil_div(T, TA_BN) * TA_BN;
  const size_t smem = static_cast<size_t>(2 * kpad) * TA_LD * 2;
  const float ik = drop_inv_keep(dc);
  attn_train_bwd_dq_kernel<<<B * nh, TA_THREADS, smem, s>>>(qkv, y, dy, lse2, T, nh, n_unmasked, dqkv, delta, dc, stream_id, ik);
  MGV_CHECK_CUDA(cudaGetLastError());
  attn_train_bwd_dkv_kernel<<<B * nh, TA_THREADS, smem + 2 * kpad * sizeof(float), s>>>(qkv, dy, lse2, delta, T, nh, n_unmasked,
                                                                                         dqkv, dc, stream_id, ik);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int train_adamw(float* p, const float* g, float* m, float* v, const AdamSeg* d_segs, const int2* d_chunks, int n_chunks,
                float lr, float beta1, float beta2, float eps, float wd, long long step, float grad_scale, cudaStream_t s) {
  MGV_REQUIRE(p && g && m && v && d_segs && d_chunks && n_chunks >= 1 && step >= 1, "adamw: bad arguments");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  adamw_kernel<<<n_chunks, 256, 0, s>>>(p, g, m, v, d_segs, d_chunks, lr, beta1, beta2, eps, wd, static_cast<float>(bc1),
                                        static_cast<float>(sqrt(bc2)), grad_scale);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

}  // namespace mgv
