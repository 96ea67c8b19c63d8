// Decode-step GEMMs whose activation operand is produced in the kernel instead of being read by TMA:
//   MODE_GELU : act = gelu_erf(src)                      (FC2 consumes the fp32 FC1 accumulator directly)
//   MODE_LN   : act = LayerNorm(src) * gamma + beta      (QKV / FC1 / head consume the fp32 residual stream directly)
// Both remove one kernel (one dependent stage of the per-position latency chain) per use.
//
// Same swap-AB structure as gemm_tc_kernel<BN, true>: weights [N_w, K] are the 128-row tcgen05 operand streamed by
// TMA (prefetched before griddepcontrol.wait), the <= 64 batch rows are the MMA N dimension, split-K over
// blockIdx.z, fp32 accumulator in TMEM, coalesced red.global.add.f32 epilogue.  The four epilogue warps first stage
// the activation tile: fp32 loads -> transform -> bf16 -> shared memory in the 128-byte-swizzled K-major layout the
// UMMA descriptor expects (16-byte chunk c of row r lives at chunk c ^ (r & 7)) -> fence.proxy.async -> mbarrier
// arrive.  MODE_LN needs full-row statistics although a CTA only holds a K-slice: the split-K CTAs of one feature
// tile form a thread-block CLUSTER, publish their per-row partial (sum, sum of squares) in shared memory and read
// each other's through distributed shared memory after one cluster barrier (fixed order: deterministic).
#include "gemm_tc.cuh"

namespace mgv {

using namespace sm100;

namespace {

constexpr int BM = 128, BK = 64, BNF = 64;             // weight rows per tile, K per stage, batch rows (MMA N)
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNF * BK * 2, STAGE = A_BYTES + B_BYTES;
constexpr int MAX_KB = 4;                               // k-blocks per CTA (all resident: no ring reuse)
constexpr int THREADS = 192;

struct FParams {
  int Nw, B, K;            // weight rows (output features), batch rows, full K
  int kb_per_split;        // 64-wide k-blocks per blockIdx.z (<= MAX_KB)
  const float* src;        // fp32 [B, K] activation source (row stride K)
  const float* gamma;      // LN
  const float* beta;       // LN
  const float* bias;       // [Nw] or null (added by split 0)
  float* out;              // fp32 [B, ldo] accumulated with atomics
  long long ldo;
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float2 ld_dsmem_f2(const float* local_ptr, uint32_t cta_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_ptr)), "r"(cta_rank));
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(raddr) : "memory");
  return v;
}

enum { MODE_GELU = 0, MODE_LN = 1 };

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_decode_fused_kernel(const __grid_constant__ CUtensorMap tmA, const FParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + MAX_KB * STAGE);   // [MAX_KB]
  uint64_t* tmem_full_bar = full_bar + MAX_KB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_part = reinterpret_cast<float*>(tmem_slot + 4);                   // [64 rows][2] partial LN statistics

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int kb0 = blockIdx.z * p.kb_per_split;
  int nkb = p.K / BK - kb0;
  if (nkb > p.kb_per_split) nkb = p.kb_per_split;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    for (int s = 0; s < MAX_KB; ++s) mbar_init(&full_bar[s], 1 + 128);   // TMA producer + 128 staging threads
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BNF);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ---- weights: independent of the upstream grid, stream them right away
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_arrive_expect_tx(&full_bar[kb], A_BYTES);
        tma_load_2d(smem + kb * STAGE, &tmA, &full_bar[kb], (kb0 + kb) * BK, m0, kEvictFirst);
      }
    }
    if (MODE == MODE_LN) {
      __syncwarp();
      cluster_arrive();
      cluster_wait();
      cluster_arrive();
    }
  } else if (warp == 1) {
    if (MODE == MODE_LN) {
      cluster_arrive();
      cluster_wait();
      cluster_arrive();
    }
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BNF);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[kb], 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + kb * STAGE);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, make_smem_desc_sw128(a_addr + k * 32), make_smem_desc_sw128(b_addr + k * 32), idesc,
                    (kb | k) != 0 ? 1u : 0u);
      }
      tc_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    // ---- activation staging by the 128 epilogue threads, then the epilogue
    pdl_wait();
    const int et = threadIdx.x - 64;        // 0..127
    const int c = et & 7;                   // 16-byte chunk (8 bf16) within the 64-wide k-block
    const int rb = et >> 3;                 // rows rb, rb+16, rb+32, rb+48
    float v[2][4][8];                       // [k-block (<=2 per pass)][row j][8 values]
    // GELU mode walks up to MAX_KB k-blocks two at a time; LN mode holds its whole slice (<= 2 k-blocks)
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, rs[4] = {1.f, 1.f, 1.f, 1.f};
    for (int kbase = 0; kbase < nkb; kbase += 2) {
      const int nk = (nkb - kbase < 2) ? nkb - kbase : 2;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = rb + 16 * j;
          float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
          if (q < nk && r < p.B) {
            const float4* s4 = reinterpret_cast<const float4*>(p.src + static_cast<long long>(r) * p.K +
                                                               (kb0 + kbase + q) * BK + c * 8);
            lo = __ldcg(s4);
            hi = __ldcg(s4 + 1);
          }
          v[q][j][0] = lo.x; v[q][j][1] = lo.y; v[q][j][2] = lo.z; v[q][j][3] = lo.w;
          v[q][j][4] = hi.x; v[q][j][5] = hi.y; v[q][j][6] = hi.z; v[q][j][7] = hi.w;
        }
      }
      if (MODE == MODE_LN) {
        // partial statistics of this CTA's K-slice, per row; 8 lanes (chunks) share a row
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float s = 0.f, ss = 0.f;
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              s += v[q][j][e];
              ss = fmaf(v[q][j][e], v[q][j][e], ss);
            }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          ss += __shfl_xor_sync(0xffffffffu, ss, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          ss += __shfl_xor_sync(0xffffffffu, ss, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          ss += __shfl_xor_sync(0xffffffffu, ss, 4);
          if (c == 0) {
            s_part[(rb + 16 * j) * 2] = s;
            s_part[(rb + 16 * j) * 2 + 1] = ss;
          }
        }
        // every CTA of the cluster (= all K-splits of this feature tile) has published its partials
        cluster_arrive();
        cluster_wait();
        const uint32_t nsplit = gridDim.z;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float s = 0.f, ss = 0.f;
          for (uint32_t z = 0; z < nsplit; ++z) {           // fixed order: deterministic
            const float2 t = ld_dsmem_f2(s_part + (rb + 16 * j) * 2, z);
            s += t.x;
            ss += t.y;
          }
          const float mean = s / static_cast<float>(p.K);
          const float var = fmaxf(ss / static_cast<float>(p.K) - mean * mean, 0.f);
          mu[j] = mean;
          rs[j] = rsqrtf(var + 1e-5f);
        }
        cluster_arrive();   // done reading remote shared memory (matched by the wait before exit)
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q < nk) {
          const int kb = kbase + q;
          float g8[8], b8[8];
          if (MODE == MODE_LN) {
            const float4* g4 = reinterpret_cast<const float4*>(p.gamma + (kb0 + kb) * BK + c * 8);
            const float4* b4 = reinterpret_cast<const float4*>(p.beta + (kb0 + kb) * BK + c * 8);
            const float4 ga = __ldg(g4), gb = __ldg(g4 + 1), ba = __ldg(b4), bb = __ldg(b4 + 1);
            g8[0] = ga.x; g8[1] = ga.y; g8[2] = ga.z; g8[3] = ga.w; g8[4] = gb.x; g8[5] = gb.y; g8[6] = gb.z; g8[7] = gb.w;
            b8[0] = ba.x; b8[1] = ba.y; b8[2] = ba.z; b8[3] = ba.w; b8[4] = bb.x; b8[5] = bb.y; b8[6] = bb.z; b8[7] = bb.w;
          }
          uint8_t* btile = smem + kb * STAGE + A_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = rb + 16 * j;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (MODE == MODE_LN) o[e] = (v[q][j][e] - mu[j]) * rs[j] * g8[e] + b8[e];
              else o[e] = gelu_erf(v[q][j][e]);
              if (r >= p.B) o[e] = 0.f;
            }
            const uint4 pk = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                        pack_bf16x2(o[6], o[7]));
            *reinterpret_cast<uint4*>(btile + r * 128 + ((c ^ (r & 7)) << 4)) = pk;   // 128-byte swizzle
          }
          fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
          mbar_arrive(&full_bar[kb]);
        }
      }
    }

    // ---- epilogue: transposed reduction, this thread owns output feature m0 + quarter*32 + lane
    const int quarter = warp & 3;
    const int feat = m0 + quarter * 32 + lane;
    const bool feat_ok = feat < p.Nw;
    const float bval = (p.bias != nullptr && blockIdx.z == 0 && feat_ok) ? __ldg(p.bias + feat) : 0.f;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float* outf = p.out + feat;
#pragma unroll 1
    for (int ch = 0; ch < BNF / 32; ++ch) {
      if (ch * 32 >= p.B) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (ch * 32 + j < p.B && feat_ok) atomicAdd(outf + static_cast<long long>(ch * 32 + j) * p.ldo, __uint_as_float(r[j]) + bval);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (MODE == MODE_LN) cluster_wait();   // no CTA of the cluster may exit while its shared memory can still be read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BNF);
  }
}

template <int MODE>
int launch_fused(const void* W, int Nw, int K, const float* src, int B, const float* gamma, const float* beta,
                 const float* bias, float* out, long long ldo, int split_k, bool pdl, cudaStream_t stream) {
  MGV_REQUIRE(B >= 1 && B <= BNF, "fused decode gemm: batch %d > %d", B, BNF);
  MGV_REQUIRE(K % BK == 0 && Nw >= 1, "fused decode gemm: bad shape");
  const int total_kb = K / BK;
  const int kbps = ceil_div(total_kb, split_k);
  const int splits = ceil_div(total_kb, kbps);
  MGV_REQUIRE(kbps <= MAX_KB, "fused decode gemm: %d k-blocks per split exceed %d", kbps, MAX_KB);
  if (MODE == MODE_LN) {
    MGV_REQUIRE(kbps <= 2, "fused LN gemm: the K slice of a CTA must fit two k-blocks (got %d)", kbps);
    MGV_REQUIRE(splits <= 8 && splits * kbps == total_kb, "fused LN gemm: splits=%d must divide K and fit a portable cluster", splits);
  }
  FParams p;
  p.Nw = Nw; p.B = B; p.K = K; p.kb_per_split = kbps;
  p.src = src; p.gamma = gamma; p.beta = beta; p.bias = bias; p.out = out; p.ldo = ldo;
  CUtensorMap tmA;
  MGV_TRY(make_tmap_2d_bf16(&tmA, W, K, Nw, static_cast<uint64_t>(K) * 2, BK, BM));
  const size_t smem = static_cast<size_t>(MAX_KB) * STAGE + (MAX_KB + 1) * 8 + 16 + 64 * 2 * 4 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_decode_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_decode_fused_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attrs[2];
  cfg.gridDim = dim3(ceil_div(Nw, BM), 1, splits);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attrs;
  cfg.numAttrs = 0;
  if (pdl) {
    attrs[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs++;
  }
  if (MODE == MODE_LN) {
    attrs[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
    attrs[cfg.numAttrs].val.clusterDim.x = 1;
    attrs[cfg.numAttrs].val.clusterDim.y = 1;
    attrs[cfg.numAttrs].val.clusterDim.z = splits;
    cfg.numAttrs++;
  }
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_decode_fused_kernel<MODE>, tmA, p));
  return MGV_OK;
}

}  // namespace

int gemm_decode_gelu(const void* W, int Nw, int K, const float* src_f32, int B, const float* bias, float* out, long long ldo,
                     int split_k, bool pdl, cudaStream_t stream) {
  return launch_fused<MODE_GELU>(W, Nw, K, src_f32, B, nullptr, nullptr, bias, out, ldo, split_k, pdl, stream);
}

int gemm_decode_ln(const void* W, int Nw, int K, const float* x_f32, int B, const float* gamma, const float* beta,
                   const float* bias, float* out, long long ldo, int split_k, bool pdl, cudaStream_t stream) {
  return launch_fused<MODE_LN>(W, Nw, K, x_f32, B, gamma, beta, bias, out, ldo, split_k, pdl, stream);
}

}  // namespace mgv
