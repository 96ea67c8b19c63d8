// Decode-step GEMMs that absorb the stage in front of them (gemm_tc.cuh: gemm_decode_fold, gpt_fold_prepare).
//
// One decode position is a chain of dependent all-to-all stages of ~4 us each, so the step gets
// shorter by REMOVING stages, not by tuning them.  Two of the seven stages of a transformer block are LayerNorms and the
// FC1 stage carried a GELU epilogue that forced it to own its full K.  Here the consumer of each of them does the work
// while it stages its activation operand (plain fp32 loads -> transform -> bf16 -> shared memory in the 128-byte
// swizzled K-major layout the UMMA descriptor expects), so a block is QKV -> attention -> proj -> FC1 -> FC2: 5 stages.
//
//   FOLD_LN   (QKV, FC1, head):  W LN(x) = rstd * (W (gamma.x) - mu * sw) + bp   with  sw = W gamma,  bp = W beta + b
//       (reference: Block.forward transformer/minGPT.py:107-119, ln_f + head :186-188).  The GEMM accumulates the RAW
//       W bf16(gamma.x) with split-K reductions; the row statistics (sum x, sum x^2 over this CTA's K slice) are written
//       as per-slice partials by the CTAs of feature tile 0 (plain stores, folded in a fixed order by the consumer:
//       deterministic, nothing to clear).  mu / rstd are applied by whoever reads the accumulator: the attention
//       kernel (q, k, v), the FOLD_GELU GEMM (FC1 output) or the sampler (logits).  sw / bp come from gpt_fold_prepare.
//   FOLD_GELU (FC2):  act = gelu_erf(rstd * (acc - mu * sw) + bp) of the raw FC1 accumulator (mlp: minGPT.py:100-105).
//
// Same swap-AB shape as gemm_tc_kernel<32, true>: 128 weight rows (MMA M) x 32 sequences (MMA N) x a K slice of at most
// four 64-wide k-blocks, weights by TMA before the grid dependency resolves, fp32 accumulator in TMEM, coalesced
// red.global.add.f32 epilogue (lane = feature).
#include <stdlib.h>
#include "gemm_tc.cuh"

namespace mgv {

using namespace sm100;

namespace {

constexpr int FBM = 128, FBK = 64;
constexpr int FA_BYTES = FBM * FBK * 2;
constexpr int F_MAX_KB = 4;

struct FoldParams {
  int Nw, B, K;              // output features (weight rows), sequences, full K
  int kbps;                  // k-blocks per split (<= F_MAX_KB)
  const float* src;          // fp32 [B, K]
  const float* gamma;        // FOLD_LN: [K]
  float2* stats_out;         // FOLD_LN: [splits][stats_stride] partial (sum, sum of squares) of the src rows
  LnFold in;                 // FOLD_GELU: statistics + fold vectors of the GEMM that produced src
  int stats_stride;          // FOLD_LN: sequences between two slices of stats_out
  const float* bias;         // [Nw] added by split 0 (or null)
  float* out;                // fp32 [B, ldo], accumulated with red.add
  long long ldo;
};

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// SW = staging warps (4 or 8), FBN = sequences per CTA (32 or 64; MMA N)
template <int MODE, int SW, int FBN>
__global__ void __launch_bounds__(64 + 32 * SW, 1)
gemm_decode_fold_kernel(const __grid_constant__ CUtensorMap tmW, const FoldParams p) {
  constexpr int FB_BYTES = FBN * FBK * 2, FSTAGE = FA_BYTES + FB_BYTES;
  constexpr int RPT = FBN / (4 * SW);  // rows per staging thread
  constexpr int RSTEP = 4 * SW;        // row distance between a thread's rows
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + F_MAX_KB * FSTAGE);
  uint64_t* x_bar = w_bar + 1;
  uint64_t* tmem_full_bar = w_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 3);
  float* s_c0 = reinterpret_cast<float*>(smem + F_MAX_KB * FSTAGE + 64);    // [F_MAX_KB * 64] gamma (FOLD_LN) / sw (FOLD_GELU)
  float* s_c1 = s_c0 + F_MAX_KB * FBK;                                      // [F_MAX_KB * 64] bp (FOLD_GELU)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * FBM;
  const int n0 = blockIdx.y * FBN;
  const int kb0 = blockIdx.z * p.kbps;
  int nkb = p.K / FBK - kb0;
  if (nkb > p.kbps) nkb = p.kbps;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmW);
    mbar_init(w_bar, 1);
    mbar_init(x_bar, 32 * SW);
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, FBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ---- weights do not depend on the upstream grid: stream them right away
    if (lane == 0) {
      mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(nkb) * FA_BYTES);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem + kb * FSTAGE, &tmW, w_bar, (kb0 + kb) * FBK, m0, kEvictFirst);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(FBM, FBN);
      mbar_wait(w_bar, 0);
      mbar_wait(x_bar, 0);
      tc_fence_after();
      const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem));
      const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem) + FA_BYTES);
      for (int kb = 0; kb < nkb; ++kb) {
        const uint64_t off = static_cast<uint64_t>((kb * FSTAGE) >> 4);
#pragma unroll
        for (int k = 0; k < FBK / 16; ++k) umma_bf16(tmem_base, da0 + off + 2 * k, db0 + off + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
      }
      tc_commit(tmem_full_bar);
    }
  } else {
    // ---- activation staging: thread (row group, 16-byte chunk c) handles 8 consecutive k of RPT rows in every k-block
    const int et = threadIdx.x - 64;
    const int c = et & 7;
    const int rb = et >> 3;
    const int kcol = kb0 * FBK + c * 8;          // first k of this thread in k-block 0
    // constants first (weights-like: safe before the dependency resolves).  They go to shared memory, not registers:
    // with 64 more live registers per thread only one CTA fits an SM and a 256-CTA grid takes two waves.
    for (int i = et; i < nkb * FBK; i += 32 * SW) {
      const float c0 = __ldg((MODE == FOLD_LN ? p.gamma : p.in.sw) + kb0 * FBK + i);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(s_c0 + i)), "f"(c0) : "memory");
      if (MODE == FOLD_GELU) {
        const float c1 = __ldg(p.in.bp + kb0 * FBK + i);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(s_c1 + i)), "f"(c1) : "memory");
      }
    }
    pdl_wait();
    float v[RPT][F_MAX_KB][8];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int r = n0 + rb + RSTEP * j;
#pragma unroll
      for (int kb = 0; kb < F_MAX_KB; ++kb) {
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
        if (kb < nkb && r < p.B) {
          const float4* s4 = reinterpret_cast<const float4*>(p.src + static_cast<long long>(r) * p.K + kcol + kb * FBK);
          lo = __ldcg(s4);
          hi = __ldcg(s4 + 1);
        }
        v[j][kb][0] = lo.x; v[j][kb][1] = lo.y; v[j][kb][2] = lo.z; v[j][kb][3] = lo.w;
        v[j][kb][4] = hi.x; v[j][kb][5] = hi.y; v[j][kb][6] = hi.z; v[j][kb][7] = hi.w;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * SW) : "memory");   // constants of all staging threads are in place
    float mu[RPT], rs[RPT];
    if (MODE == FOLD_LN) {
      // partial statistics of the RAW rows over this K slice (the 8 chunk lanes of a row are consecutive lanes)
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int kb = 0; kb < F_MAX_KB; ++kb)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            s += v[j][kb][e];
            ss = fmaf(v[j][kb][e], v[j][kb][e], ss);
          }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        const int r = n0 + rb + RSTEP * j;
        if (blockIdx.x == 0 && c == 0 && r < p.B)
          p.stats_out[static_cast<long long>(blockIdx.z) * p.stats_stride + r] = make_float2(s, ss);
        mu[j] = 0.f;
        rs[j] = 1.f;
      }
    } else {
      // statistics of the producer's LayerNorm: chunk lane c folds slices c, c+8, ...; fixed order -> deterministic
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int r = n0 + rb + RSTEP * j;
        float s = 0.f, ss = 0.f;
        if (r < p.B)
          for (int z = c; z < p.in.nparts; z += 8) {
            const float2 t = __ldcg(p.in.stats + static_cast<long long>(z) * p.in.stride + r);
            s += t.x;
            ss += t.y;
          }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        const float inv_n = 1.0f / static_cast<float>(p.in.dim);
        const float mean = s * inv_n;
        mu[j] = mean;
        rs[j] = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + 1e-5f);
      }
    }
#pragma unroll
    for (int kb = 0; kb < F_MAX_KB; ++kb) {
      if (kb < nkb) {
        uint8_t* btile = smem + kb * FSTAGE + FA_BYTES;
        float cg[8], cb[8];
        {
          const float4 a = *reinterpret_cast<const float4*>(s_c0 + kb * FBK + c * 8);
          const float4 b = *reinterpret_cast<const float4*>(s_c0 + kb * FBK + c * 8 + 4);
          cg[0] = a.x; cg[1] = a.y; cg[2] = a.z; cg[3] = a.w; cg[4] = b.x; cg[5] = b.y; cg[6] = b.z; cg[7] = b.w;
          if (MODE == FOLD_GELU) {
            const float4 d = *reinterpret_cast<const float4*>(s_c1 + kb * FBK + c * 8);
            const float4 e = *reinterpret_cast<const float4*>(s_c1 + kb * FBK + c * 8 + 4);
            cb[0] = d.x; cb[1] = d.y; cb[2] = d.z; cb[3] = d.w; cb[4] = e.x; cb[5] = e.y; cb[6] = e.z; cb[7] = e.w;
          }
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int rl = rb + RSTEP * j;
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (MODE == FOLD_LN) o[e] = v[j][kb][e] * cg[e];
            else {
              const float pre = fmaf(rs[j], fmaf(-mu[j], cg[e], v[j][kb][e]), cb[e]);
              o[e] = gelu_erf_fast(pre);
            }
            if (n0 + rl >= p.B) o[e] = 0.f;
          }
          const uint4 pk = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                      pack_bf16x2(o[6], o[7]));
          *reinterpret_cast<uint4*>(btile + rl * 128 + ((c ^ (rl & 7)) << 4)) = pk;   // 128-byte swizzle
        }
      }
    }
    fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    mbar_arrive(x_bar);

    // ---- epilogue (warps 2..5: one TMEM lane quarter each): lane = output feature, columns = sequences
    if (warp < 6) {
      const int quarter = warp & 3;
      const int feat = m0 + quarter * 32 + lane;
      const bool feat_ok = feat < p.Nw;
      const float bval = (p.bias != nullptr && blockIdx.z == 0 && feat_ok) ? __ldg(p.bias + feat) : 0.f;
      float* outf = p.out + static_cast<long long>(n0) * p.ldo + feat;
      const int ncols = (p.B - n0 < FBN) ? p.B - n0 : FBN;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < FBN / 32; ++ch) {
        if (ch * 32 >= ncols) break;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ch * 32, r);
        tmem_ld_wait();
        if (feat_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ch * 32 + j < ncols) red_add_f32(outf + static_cast<long long>(ch * 32 + j) * p.ldo, __uint_as_float(r[j]) + bval);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FBN);
  }
}

template <int MODE, int SW, int FBN>
int launch_fold(const CUtensorMap& tmW, const FoldParams& p, int splits, bool pdl, cudaStream_t stream) {
  constexpr int FSTAGE = FA_BYTES + FBN * FBK * 2;
  const size_t smem = static_cast<size_t>(F_MAX_KB) * FSTAGE + 64 + 2 * F_MAX_KB * FBK * 4 + 1024;
  static unsigned long long attr_mask = 0;   // per device (and per template instantiation)
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_decode_fold_kernel<MODE, SW, FBN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_decode_fold_kernel<MODE, SW, FBN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
  }
  LaunchCfg lc(dim3(ceil_div(p.Nw, FBM), ceil_div(p.B, FBN), splits), dim3(64 + 32 * SW), smem, stream, pdl);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, gemm_decode_fold_kernel<MODE, SW, FBN>, tmW, p));
  return MGV_OK;
}

// sw[n] = sum_k W[n,k] gamma[k] ; bp[n] = sum_k W[n,k] beta[k] + bias[n]     (one warp per weight row)
__global__ void __launch_bounds__(128)
fold_prepare_kernel(const __nv_bfloat16* __restrict__ W, int N, int K, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ bias, float* __restrict__ sw,
                    float* __restrict__ bp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + warp;
  if (n >= N) return;
  const __nv_bfloat16* row = W + static_cast<long long>(n) * K;
  float s = 0.f, t = 0.f;
  for (int k = lane * 2; k < K; k += 64) {
    const float2 w = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + k));
    s = fmaf(w.x, gamma[k], s);
    s = fmaf(w.y, gamma[k + 1], s);
    t = fmaf(w.x, beta[k], t);
    t = fmaf(w.y, beta[k + 1], t);
  }
  s = warp_sum(s);
  t = warp_sum(t);
  if (lane == 0) {
    sw[n] = s;
    bp[n] = t + (bias ? bias[n] : 0.f);
  }
}

// out[b, n] = rstd_b * (out[b, n] - mu_b * sw[n]) + bp[n]   (tests: what the consumers of a FOLD_LN accumulator do)
__global__ void fold_apply_kernel(float* __restrict__ out, int B, int N, const LnFold f) {
  const int b = blockIdx.x;
  float s = 0.f, ss = 0.f;
  for (int z = 0; z < f.nparts; ++z) {
    const float2 p = f.stats[static_cast<long long>(z) * f.stride + b];
    s += p.x;
    ss += p.y;
  }
  const float inv_n = 1.0f / static_cast<float>(f.dim);
  const float mean = s * inv_n;
  const float rstd = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + 1e-5f);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float* o = out + static_cast<long long>(b) * N + n;
    *o = fmaf(rstd, fmaf(-mean, f.sw[n], *o), f.bp[n]);
  }
}

}  // namespace

int gpt_fold_apply(float* out, int B, int N, const LnFold& f, cudaStream_t stream) {
  MGV_REQUIRE(out && f.stats && f.sw && f.bp && B >= 1 && N >= 1, "fold apply: bad arguments");
  fold_apply_kernel<<<B, 256, 0, stream>>>(out, B, N, f);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int gemm_decode_fold(int mode, const void* W, int Nw, int K, const float* src, int B, const float* gamma,
                     float2* stats_out, int stats_stride, const LnFold* in, const float* bias, float* out, long long ldo,
                     int kbps, int staging_warps, int bn, bool pdl, cudaStream_t stream) {
  MGV_REQUIRE(W && src && out && Nw >= 1 && B >= 1, "fold decode gemm: bad arguments");
  MGV_REQUIRE(K % FBK == 0 && kbps >= 1 && kbps <= F_MAX_KB, "fold decode gemm: K=%d kbps=%d", K, kbps);
  const int splits = ceil_div(K / FBK, kbps);
  FoldParams p;
  memset(&p, 0, sizeof(p));
  p.Nw = Nw; p.B = B; p.K = K; p.kbps = kbps;
  p.src = src; p.bias = bias; p.out = out; p.ldo = ldo;
  CUtensorMap tmW;
  MGV_TRY(make_tmap_2d_bf16(&tmW, W, K, Nw, static_cast<uint64_t>(K) * 2, FBK, FBM));
  const bool wide = staging_warps >= 8 || bn == 64;
  MGV_REQUIRE(bn == 32 || bn == 64, "fold decode gemm: bn=%d", bn);
  if (mode == FOLD_LN) {
    MGV_REQUIRE(gamma && stats_out && stats_stride >= B, "fold decode gemm: LN mode needs gamma and a statistics buffer");
    p.gamma = gamma; p.stats_out = stats_out; p.stats_stride = stats_stride;
    if (bn == 64) return launch_fold<FOLD_LN, 8, 64>(tmW, p, splits, pdl, stream);
    return wide ? launch_fold<FOLD_LN, 8, 32>(tmW, p, splits, pdl, stream) : launch_fold<FOLD_LN, 4, 32>(tmW, p, splits, pdl, stream);
  }
  MGV_REQUIRE(mode == FOLD_GELU && in && in->stats && in->sw && in->bp && in->nparts >= 1 && in->dim >= 1,
              "fold decode gemm: GELU mode needs the producer's statistics and fold vectors");
  p.in = *in;
  if (bn == 64) return launch_fold<FOLD_GELU, 8, 64>(tmW, p, splits, pdl, stream);
  return wide ? launch_fold<FOLD_GELU, 8, 32>(tmW, p, splits, pdl, stream) : launch_fold<FOLD_GELU, 4, 32>(tmW, p, splits, pdl, stream);
}

int gpt_fold_prepare(const void* W, int N, int K, const float* gamma, const float* beta, const float* bias, float* sw,
                     float* bp, cudaStream_t stream) {
  MGV_REQUIRE(W && gamma && beta && sw && bp && N >= 1 && K % 2 == 0, "fold prepare: bad arguments");
  fold_prepare_kernel<<<ceil_div(N, 4), 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(W), N, K, gamma, beta, bias, sw, bp);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

}  // namespace mgv
