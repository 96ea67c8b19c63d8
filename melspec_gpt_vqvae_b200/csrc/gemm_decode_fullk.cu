// Decode-step GEMM whose CTAs own their full K extent: out[b, n] = epi(sum_k X[b, k] * W[n, k] + bias[n]).
//
// A CTA = 128 weight rows (MMA M) x 32 sequences (MMA N) x all of K, so there is no split-K reduction and the
// epilogue can apply bias + GELU and store bf16 directly: FC1 needs no separate GELU stage (one all-to-all stage less
// per transformer block; a stage costs 4-5 us of pure latency in the decode chain).  The price is 2 * K * 128 bytes of
// weights per CTA (256 KB at K = 1024), more than shared memory holds, so the k-blocks travel in GROUPS of four through
// a two-group ring: one barrier hand-off per 16 MMAs (a hand-off costs the issuing threads ~0.25 us, see
// gemm_tc_persist_kernel), and the first two groups of weights (half of them) are in flight before the grid
// dependency resolves.
#include "gemm_tc.cuh"
#include "mgv_sm100.cuh"
#include <stdlib.h>

namespace mgv {

using namespace sm100;

namespace {

constexpr int FK_THREADS = 192;
constexpr int FK_BM = 128, FK_BN = 32, FK_BK = 64;
constexpr int FK_A_BYTES = FK_BM * FK_BK * 2, FK_B_BYTES = FK_BN * FK_BK * 2;
constexpr int FK_KB_BYTES = FK_A_BYTES + FK_B_BYTES;           // one k-block: weight tile, then activation tile
constexpr int FK_GROUPS = 2;                                   // ring depth

struct FkParams {
  int Nw, B, K, epi;
  const float* bias;
  void* out;
  const void* resid;
  long long ldo;
};

template <int FK_KG>   // k-blocks per group (4: 80 KB groups; 2: 40 KB groups, room for the next kernel's CTAs on the SM)
__global__ void __launch_bounds__(FK_THREADS, 1)
gemm_decode_fullk_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const FkParams p) {
  constexpr int FK_GROUP_BYTES = FK_KG * FK_KB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + FK_GROUPS * FK_GROUP_BYTES);   // [2]
  uint64_t* empty_bar = full_bar + FK_GROUPS;                                             // [2]
  uint64_t* tmem_full_bar = empty_bar + FK_GROUPS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * FK_BM;       // first weight row (output feature)
  const int n0 = blockIdx.y * FK_BN;       // first sequence
  const int ng = p.K / (FK_BK * FK_KG);    // groups

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmX);
    for (int g = 0; g < FK_GROUPS; ++g) {
      mbar_init(&full_bar[g], 1);
      mbar_init(&empty_bar[g], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, FK_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      auto load_w = [&](int g) {
        uint8_t* dst = smem + (g & 1) * FK_GROUP_BYTES;
        for (int j = 0; j < FK_KG; ++j)
          tma_load_2d(dst + j * FK_KB_BYTES, &tmW, &full_bar[g & 1], (g * FK_KG + j) * FK_BK, m0, kEvictNormal);
      };
      auto load_x = [&](int g) {
        uint8_t* dst = smem + (g & 1) * FK_GROUP_BYTES + FK_A_BYTES;
        for (int j = 0; j < FK_KG; ++j)
          tma_load_2d(dst + j * FK_KB_BYTES, &tmX, &full_bar[g & 1], (g * FK_KG + j) * FK_BK, n0, kEvictNormal);
      };
      const int pre = ng < FK_GROUPS ? ng : FK_GROUPS;
      for (int g = 0; g < pre; ++g) {       // weights do not depend on the upstream grid
        mbar_arrive_expect_tx(&full_bar[g], FK_GROUP_BYTES);
        load_w(g);
      }
      // the weight tiles that do not fit the ring yet: pull them into L2 now, so that their turn costs an L2 hit
      // instead of an HBM round trip on the critical path
      for (int kb = pre * FK_KG; kb < ng * FK_KG; ++kb)
        asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(&tmW)),
                     "r"(kb * FK_BK), "r"(m0)
                     : "memory");
      pdl_wait();
      for (int g = 0; g < pre; ++g) load_x(g);
      for (int g = pre; g < ng; ++g) {
        mbar_wait(&empty_bar[g & 1], ((g >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[g & 1], FK_GROUP_BYTES);
        load_w(g);
        load_x(g);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(FK_BM, FK_BN);
      const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem));
      const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem) + FK_A_BYTES);
      uint32_t acc = 0;
      for (int g = 0; g < ng; ++g) {
        mbar_wait(&full_bar[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const uint64_t goff = static_cast<uint64_t>(((g & 1) * FK_GROUP_BYTES) >> 4);
#pragma unroll
        for (int j = 0; j < FK_KG; ++j) {
#pragma unroll
          for (int k = 0; k < FK_BK / 16; ++k) {
            const uint64_t off = goff + static_cast<uint64_t>((j * FK_KB_BYTES) >> 4) + 2 * k;
            umma_bf16(tmem_base, da0 + off, db0 + off, idesc, acc);
            acc = 1;
          }
        }
        if (g + FK_GROUPS < ng) tc_commit(&empty_bar[g & 1]);   // the slot is refilled only if a later group needs it
      }
      tc_commit(tmem_full_bar);
    }
  } else {
    pdl_wait();   // residual / output buffers belong to the upstream grid
    const int quarter = warp & 3;
    const int feat = m0 + quarter * 32 + lane;
    const bool feat_ok = feat < p.Nw;
    const float bval = (p.bias != nullptr && feat_ok) ? __ldg(p.bias + feat) : 0.f;
    const int ncols = (p.B - n0 < FK_BN) ? p.B - n0 : FK_BN;
    float res[FK_BN];
    if (p.epi == EPI_F32_RESID && feat_ok) {   // residual loads overlap the MMAs
#pragma unroll
      for (int j = 0; j < FK_BN; ++j)
        if (j < ncols) res[j] = static_cast<const float*>(p.resid)[static_cast<long long>(n0 + j) * p.ldo + feat];
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16), r);
    tmem_ld_wait();
    if (feat_ok) {
      // for a fixed sequence the 32 lanes of a warp write 32 consecutive features: coalesced
      if (p.epi == EPI_BF16_GELU) {
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(n0) * p.ldo + feat;
#pragma unroll
        for (int j = 0; j < FK_BN; ++j)
          if (j < ncols) o[j * p.ldo] = __float2bfloat16(gelu_erf(__uint_as_float(r[j]) + bval));
      } else {
        float* o = static_cast<float*>(p.out) + static_cast<long long>(n0) * p.ldo + feat;
#pragma unroll
        for (int j = 0; j < FK_BN; ++j)
          if (j < ncols) o[j * p.ldo] = __uint_as_float(r[j]) + bval + (p.epi == EPI_F32_RESID ? res[j] : 0.f);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FK_BN);
  }
}

template <int KG>
int launch_fullk(const CUtensorMap& tmW, const CUtensorMap& tmX, const FkParams& p, bool pdl, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(FK_GROUPS) * KG * FK_KB_BYTES + (2 * FK_GROUPS + 1) * 8 + 16 + 1024;
  static unsigned long long attr_mask = 0;   // per device (and per template instantiation)
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_decode_fullk_kernel<KG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_decode_fullk_kernel<KG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
  }
  LaunchCfg lc(dim3(ceil_div(p.Nw, FK_BM), ceil_div(p.B, FK_BN)), dim3(FK_THREADS), smem, stream, pdl);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, gemm_decode_fullk_kernel<KG>, tmW, tmX, p));
  return MGV_OK;
}

}  // namespace

int gemm_decode_fullk(const void* W, int Nw, int K, const void* X_bf16, int B, const float* bias, int epi, void* out,
                      const void* resid, long long ldo, bool pdl, cudaStream_t stream) {
  MGV_REQUIRE(W && X_bf16 && out && Nw >= 1 && B >= 1, "full-K decode gemm: bad arguments");
  MGV_REQUIRE(K % (FK_BK * 4) == 0, "full-K decode gemm: K=%d must be a multiple of %d", K, FK_BK * 4);
  MGV_REQUIRE(epi == EPI_BF16_GELU || epi == EPI_F32 || epi == EPI_F32_RESID, "full-K decode gemm: epilogue %d unsupported", epi);
  MGV_REQUIRE(epi != EPI_F32_RESID || resid != nullptr, "full-K decode gemm: residual epilogue without a residual");
  FkParams p;
  p.Nw = Nw; p.B = B; p.K = K; p.epi = epi; p.bias = bias; p.out = out; p.resid = resid; p.ldo = ldo;
  CUtensorMap tmW, tmX;
  MGV_TRY(make_tmap_2d_bf16(&tmW, W, K, Nw, static_cast<uint64_t>(K) * 2, FK_BK, FK_BM));
  MGV_TRY(make_tmap_2d_bf16(&tmX, X_bf16, K, B, static_cast<uint64_t>(K) * 2, FK_BK, FK_BN));
  static const int kg = getenv("MGV_FK_KG") ? atoi(getenv("MGV_FK_KG")) : 4;
  return kg == 2 ? launch_fullk<2>(tmW, tmX, p, pdl, stream) : launch_fullk<4>(tmW, tmX, p, pdl, stream);
}

}  // namespace mgv
