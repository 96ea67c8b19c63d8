// Persistent stage-program kernel of the decode step (see gpt_decode_program.cuh).
//
// Why: at batch 64 one decode position is a chain of ~190 dependent kernels of 3-5 us each.  Every stage is an
// all-to-all dependency (each GEMM tile needs every row of the previous stage), so the chain cannot be shortened
// by overlapping stages -- only by making a stage cheaper and by having fewer of them.  Here the stages of one
// transformer block run inside ONE kernel with 148 resident CTAs:
//   * the weights of the next GEMM stage stream into shared memory while the current stage and the barrier are
//     in flight, so a GEMM unit may own its full K extent (<= 1024): no split-K, hence no atomic reductions
//     (64 REDs per thread cost more than everything else in the stage), and bias / GELU / residual fuse into the
//     epilogue -- the separate GELU stage disappears;
//   * tensor memory, barriers and tensor maps are set up once per block instead of once per stage;
//   * a stage boundary is one release/acquire round on a global counter.
//
// CTA = 6 warps: warp 0 = TMA producer (+ the grid-barrier spinner), warp 1 = tcgen05 issuer, warps 2-5 = epilogue /
// LayerNorm workers (TMEM lane quarter = warp & 3).
// GEMM unit = 64 weight rows (MMA M = 64) x 32 sequences (MMA N = 32) x K slice; cta_group::1 M=64 accumulators
// live in lanes 0-15 of each 32-lane TMEM quarter (row 16*q + i -> lane 32*q + i).
#include "gpt_decode_program.cuh"

namespace mgv {

using namespace sm100;

namespace {

constexpr int DP_THREADS = 192;
constexpr int DP_BM = 64;                        // weight rows per unit
constexpr int DP_BN = 32;                        // sequences per unit
constexpr int DP_W_KB_BYTES = DP_BM * 64 * 2;    // 8 KB: 64 weight rows x 64 k (bf16), 128-byte swizzle rows
constexpr int DP_X_KB_BYTES = DP_BN * 64 * 2;    // 4 KB: 32 sequences x 64 k
constexpr int DP_LN_MAX_V4 = 8;                  // C <= 1024
constexpr int DP_MAX_STAGES = 8;                 // stages per launch

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// The stage descriptors live in shared memory, so the compiler cannot tell that their pointers are global:
// spell the address space out (a generic atomicAdd becomes a synchronous ATOM with shared/local fallbacks).
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}
__device__ __forceinline__ float ldg_f32(const float* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(__cvta_generic_to_global(p)) : "memory");
  return v;
}
__device__ __forceinline__ void stg_f32(float* p, float v) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_bf16(__nv_bfloat16* p, float v) {
  const unsigned short u = __bfloat16_as_ushort(__float2bfloat16(v));
  asm volatile("st.global.u16 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "h"(u) : "memory");
}
__device__ __forceinline__ float4 ldg_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(__cvta_generic_to_global(p))
               : "memory");
  return v;
}
__device__ __forceinline__ void stg_u2(uint2* p, uint2 v) {
  asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(p)), "r"(v.x), "r"(v.y) : "memory");
}

struct Unit {
  bool has;
  int ftile, rhalf, split, kb0, n;
};
__device__ __forceinline__ Unit unit_of(const DecStage& d) {
  Unit u;
  const int id = blockIdx.x;
  const int per_split = d.ftiles * d.rhalves;
  u.has = id < per_split * d.splits;
  u.split = id / per_split;
  const int rem = id - u.split * per_split;
  u.rhalf = rem / d.ftiles;
  u.ftile = rem - u.rhalf * d.ftiles;
  u.kb0 = u.split * d.nkb / d.splits;
  u.n = (u.split + 1) * d.nkb / d.splits - u.kb0;
  return u;
}

// weights of this CTA's unit of GEMM stage d -> wbuf (one thread)
__device__ __forceinline__ void issue_weights(const DecStage& d, const CUtensorMap* maps, uint8_t* wbuf, uint64_t* w_bar) {
  const Unit u = unit_of(d);
  if (!u.has) return;
  mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(u.n) * DP_W_KB_BYTES);
  for (int k = 0; k < u.n; ++k)
    tma_load_2d(wbuf + k * DP_W_KB_BYTES, &maps[d.map_w], w_bar, (u.kb0 + k) * 64, u.ftile * DP_BM, kEvictNormal);
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) : : "memory");
  return t;
}

// all CTAs of the grid: arrive, then wait until `target` arrivals (counter is zero at kernel start)
__device__ __forceinline__ void grid_sync(unsigned int* counter, unsigned int target, unsigned long long* trace) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (trace) trace[0] = gtimer();
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 23)) __trap();   // a protocol bug becomes a CUDA error instead of a hung GPU
    } while (v < target);
    fence_proxy_async_all();   // activations written with generic stores by other CTAs are read by TMA next
    if (trace) trace[1] = gtimer();
  }
  __syncthreads();
  tc_fence_after();
}

__global__ void __launch_bounds__(DP_THREADS, 1)
decode_program_kernel(const DecStage* __restrict__ gprog, int s_begin, int s_end, const CUtensorMap* __restrict__ maps,
                      unsigned int* counter, unsigned long long* trace) {
  // this launch's stages, staged in shared memory (a global-memory read per field costs a round trip on the
  // critical path)
  __shared__ DecStage prog_s[DP_MAX_STAGES];
  const int n_stages = s_end - s_begin;
  for (int i = threadIdx.x; i < n_stages * static_cast<int>(sizeof(DecStage) / 4); i += DP_THREADS)
    reinterpret_cast<uint32_t*>(prog_s)[i] = reinterpret_cast<const uint32_t*>(gprog + s_begin)[i];
  const DecStage* prog = prog_s - s_begin;   // prog[s] for s in [s_begin, s_end)
  if (blockIdx.x != 0) trace = nullptr;
  if (trace && threadIdx.x == 0) trace[0] = gtimer();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wbuf = smem;
  uint8_t* xbuf = smem + DP_MAX_KB * DP_W_KB_BYTES;
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(xbuf + DP_MAX_KB * DP_X_KB_BYTES);
  uint64_t* x_bar = w_bar + 1;
  uint64_t* mma_bar = w_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(w_bar, 1);
    mbar_init(x_bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 4 * DP_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  // weights of the first GEMM stage do not depend on the upstream grid
  if (threadIdx.x == 0) {
    for (int t = s_begin; t < s_end; ++t)
      if (prog[t].type == DST_GEMM) {
        prefetch_tensormap(&maps[prog[t].map_w]);
        prefetch_tensormap(&maps[prog[t].map_x]);
      }
    int t = s_begin;
    while (t < s_end && prog[t].type != DST_GEMM) ++t;
    if (t < s_end) issue_weights(prog[t], maps, wbuf, w_bar);
  }
  pdl_wait();
  if (trace && threadIdx.x == 0) trace[1] = gtimer();

  int gemm_no = 0;
  uint32_t par = 0;          // phase parity of w_bar / x_bar / mma_bar: flips after every GEMM stage this CTA worked on
  unsigned int syncs = 0;
  for (int s = s_begin; s < s_end; ++s) {
    const DecStage& d = prog[s];
    if (d.type == DST_GEMM) {
      const Unit u = unit_of(d);
      if (u.has) {
        if (warp == 0) {
          if (lane == 0) {
            mbar_arrive_expect_tx(x_bar, static_cast<uint32_t>(u.n) * DP_X_KB_BYTES);
            for (int k = 0; k < u.n; ++k)
              tma_load_2d(xbuf + k * DP_X_KB_BYTES, &maps[d.map_x], x_bar, (u.kb0 + k) * 64, u.rhalf * DP_BN, kEvictNormal);
          }
        } else if (warp == 1) {
          if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(DP_BM, DP_BN);
            mbar_wait(w_bar, par);
            if (trace && gemm_no < 4) trace[16 + 4 * gemm_no + 3] = gtimer();
            mbar_wait(x_bar, par);
            if (trace && gemm_no < 4) trace[16 + 4 * gemm_no + 0] = gtimer();
            tc_fence_after();
            for (int k = 0; k < u.n; ++k) {
              const uint32_t a_addr = smem_u32(wbuf + k * DP_W_KB_BYTES);
              const uint32_t b_addr = smem_u32(xbuf + k * DP_X_KB_BYTES);
              // four independent accumulators (one per 16-wide k step of the block): a 64x32 MMA is far shorter than
              // the accumulate latency, so a single dependent chain of K/16 MMAs would run at ~40 ns per MMA
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16(tmem_base + kk * DP_BN, make_smem_desc_sw128(a_addr + kk * 32),
                          make_smem_desc_sw128(b_addr + kk * 32), idesc, k != 0 ? 1u : 0u);
            }
            tc_commit(mma_bar);
          }
        } else {
          // thread (quarter q, lane i < 16) owns weight row 16*q + i of the tile; accumulator columns = sequences
          const int quarter = warp & 3;
          const int feat = u.ftile * DP_BM + quarter * 16 + (lane & 15);
          const bool act = lane < 16 && feat < d.n_feat;
          const int b0 = u.rhalf * DP_BN;
          const int nb = min(DP_BN, d.B - b0);   // valid sequences of this unit
          const long long ldo = d.ldo;
          const int mode = d.mode;
          const float bval = (d.bias != nullptr && act && (mode != DGM_RED_F32 || u.split == 0)) ? ldg_f32(d.bias + feat) : 0.f;
          float resid[DP_BN];
          if (mode == DGM_ADD_F32 && act) {   // residual loads overlap the MMAs
            const float* o = static_cast<const float*>(d.out) + static_cast<long long>(b0) * ldo + feat;
#pragma unroll
            for (int j = 0; j < DP_BN; ++j)
              if (j < nb) resid[j] = ldg_f32(o + j * ldo);
          }
          mbar_wait(mma_bar, par);
          if (trace && threadIdx.x == 64 && gemm_no < 4) trace[16 + 4 * gemm_no + 1] = gtimer();
          tc_fence_after();
          uint32_t r[32];
          {
            uint32_t r1[32], r2[32], r3[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
            tmem_ld_32x32(taddr, r);
            tmem_ld_32x32(taddr + DP_BN, r1);
            tmem_ld_32x32(taddr + 2 * DP_BN, r2);
            tmem_ld_32x32(taddr + 3 * DP_BN, r3);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              r[j] = __float_as_uint((__uint_as_float(r[j]) + __uint_as_float(r1[j])) +
                                     (__uint_as_float(r2[j]) + __uint_as_float(r3[j])));
          }
          if (act) {
            if (mode == DGM_GELU_BF16) {
              __nv_bfloat16* o = static_cast<__nv_bfloat16*>(d.out) + static_cast<long long>(b0) * ldo + feat;
#pragma unroll
              for (int j = 0; j < DP_BN; ++j)
                if (j < nb) stg_bf16(o + j * ldo, gelu_erf(__uint_as_float(r[j]) + bval));
            } else {
              float* o = static_cast<float*>(d.out) + static_cast<long long>(b0) * ldo + feat;
              if (mode == DGM_RED_F32) {
#pragma unroll
                for (int j = 0; j < DP_BN; ++j)
                  if (j < nb) red_add_f32(o + j * ldo, __uint_as_float(r[j]) + bval);
              } else if (mode == DGM_ADD_F32) {
#pragma unroll
                for (int j = 0; j < DP_BN; ++j)
                  if (j < nb) stg_f32(o + j * ldo, resid[j] + (__uint_as_float(r[j]) + bval));
              } else {
#pragma unroll
                for (int j = 0; j < DP_BN; ++j)
                  if (j < nb) stg_f32(o + j * ldo, __uint_as_float(r[j]) + bval);
              }
            }
          }
          if (mode == DGM_GELU_BF16) fence_proxy_async_all();   // read by TMA in the next stage
        }
      }
      if (trace && threadIdx.x == 64 && gemm_no < 4) trace[16 + 4 * gemm_no + 2] = gtimer();
      // the weight buffer is free once this stage's MMAs have retired: start streaming the next GEMM's weights
      if (threadIdx.x == 0) {
        if (u.has) mbar_wait(mma_bar, par);
        int t = s + 1;
        while (t < s_end && prog[t].type != DST_GEMM) ++t;
        if (t < s_end) issue_weights(prog[t], maps, wbuf, w_bar);
      }
      if (u.has) par ^= 1;
      ++gemm_no;
    } else {  // DST_LN: one row per worker warp
      const int row = blockIdx.x * 4 + (warp - 2);
      if (warp >= 2 && row < d.B) {
        const int C = d.C, nv = C / 4;
        const float4* x4 = reinterpret_cast<const float4*>(d.x + static_cast<long long>(row) * C);
        const float4* w4 = reinterpret_cast<const float4*>(d.gamma);
        const float4* b4 = reinterpret_cast<const float4*>(d.beta);
        uint2* o2 = reinterpret_cast<uint2*>(d.ln_out + static_cast<long long>(row) * C);
        float4 v[DP_LN_MAX_V4], g[DP_LN_MAX_V4], be[DP_LN_MAX_V4];
#pragma unroll
        for (int j = 0; j < DP_LN_MAX_V4; ++j) {
          const int i = lane + 32 * j;
          if (i < nv) v[j] = ldg_f4(x4 + i);
        }
#pragma unroll
        for (int j = 0; j < DP_LN_MAX_V4; ++j) {   // parameters: in flight during the reductions
          const int i = lane + 32 * j;
          if (i < nv) {
            g[j] = ldg_f4(w4 + i);
            be[j] = ldg_f4(b4 + i);
          }
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < DP_LN_MAX_V4; ++j) {
          const int i = lane + 32 * j;
          if (i < nv) sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        }
        const float mean = warp_sum(sum) / static_cast<float>(C);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < DP_LN_MAX_V4; ++j) {
          const int i = lane + 32 * j;
          if (i < nv) {
            const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
            ss += (a * a + b * b) + (c * c + e * e);
          }
        }
        const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(C) + 1e-5f);
#pragma unroll
        for (int j = 0; j < DP_LN_MAX_V4; ++j) {
          const int i = lane + 32 * j;
          if (i < nv) {
            const float a = (v[j].x - mean) * rstd * g[j].x + be[j].x;
            const float b = (v[j].y - mean) * rstd * g[j].y + be[j].y;
            const float c = (v[j].z - mean) * rstd * g[j].z + be[j].z;
            const float e = (v[j].w - mean) * rstd * g[j].w + be[j].w;
            stg_u2(o2 + i, make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, e)));
          }
        }
        fence_proxy_async_all();   // read by TMA in the next stage
      }
    }
    if (s + 1 < s_end) {
      grid_sync(counter, (syncs + 1) * gridDim.x, trace ? trace + 2 + 2 * syncs : nullptr);
      ++syncs;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (trace && threadIdx.x == 0) trace[2 + 2 * syncs] = gtimer();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 4 * DP_BN);
  }
}

}  // namespace

int decode_program_launch(const DecStage* prog, int s_begin, int s_end, const CUtensorMap* maps, unsigned int* counter,
                          int n_ctas, bool pdl, cudaStream_t s, unsigned long long* trace) {
  MGV_REQUIRE(prog && maps && counter && s_end > s_begin && n_ctas >= 1, "decode program: bad arguments");
  MGV_REQUIRE(s_end - s_begin <= DP_MAX_STAGES, "decode program: more than %d stages per launch", DP_MAX_STAGES);
  static_assert(sizeof(DecStage) % 4 == 0, "DecStage is copied word by word");
  const size_t smem = 1024 + DP_MAX_KB * (DP_W_KB_BYTES + DP_X_KB_BYTES) + 64;
  static bool attr_set = false;
  if (!attr_set) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(decode_program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(decode_program_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  LaunchCfg lc(dim3(n_ctas), dim3(DP_THREADS), smem, s, pdl);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, decode_program_kernel, prog, s_begin, s_end, maps, counter, trace));
  return MGV_OK;
}

}  // namespace mgv
