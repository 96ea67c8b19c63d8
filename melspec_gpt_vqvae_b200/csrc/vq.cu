// Fused codebook kernels.
//   reference: VectorQuantizer.forward            vqvae/big_model_attn_gan.py:19-54
//              VectorQuantizer.get_codebook_entry vqvae/big_model_attn_gan.py:56-71
//
// vq_argmin: for every latent vector x (read straight from the BCHW tensor, no permute copy) the index of the nearest
// code in fp32 and in the reference's operation order,
//     d_j = fl( fl(|x|^2 + |e_j|^2) - fl(2 * <x, e_j>) )          (reference :28-30)
// and the first index attaining the minimum (torch.argmin semantics, :33).  Every sum is a sequential fmaf chain over
// the channel index (k = 0..D-1, starting from +0.0f); oracle/vq_oracle.c restates exactly this, so indices AND
// distances are bit-reproducible on the CPU.
//
// Two kernels.  (1) vq_prefilter_kernel streams z once from HBM (cp.async, 64-channel chunks, double buffered across
// tiles), forms all 128 approximate distances of a vector on the tensor cores (TF32 mma.sync, codebook resident in
// shared memory) and keeps as candidates the codes within a guaranteed error margin of the minimum -- typically one,
// sometimes two.  (2) vq_exact_kernel evaluates the oracle's exact chain only for the candidates (re-reading the few
// needed columns of z from L2) and takes the exact argmin among them.  The result is bit-identical to evaluating all
// 128 exact distances (the previous FMA-pipe kernel: 2.2 G fmaf, 143 us at N = 67 840) because the margin provably
// contains the exact winner and everything tied with it.
#include <stdlib.h>
#include "mgv_sm100.cuh"

namespace mgv {

namespace {

constexpr int VQ_TILE_V = 128;   // vectors per prefilter tile
constexpr int VQ_MAX_K = 128;    // codes per pass
constexpr int VQ_EXACT_TILE = 512;      // vectors per tile of the exact kernel
constexpr int VQ_EXACT_THREADS = 512;

// order-preserving float -> uint32 map (and back) for integer atomicMin on distances
__device__ __forceinline__ uint32_t f32_ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unord(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

// codebook chunk -> shared memory, row major es[code][k] with pitch SA = D + 4 (4 mod 32: A fragments conflict free);
// rows of absent codes are zero.  ee[code] = |e|^2 as the oracle's sequential fmaf chain (+inf for absent codes).
__device__ __forceinline__ void stage_codebook(const float* __restrict__ codebook, int D, int K, float* es, float* ee,
                                               int nthreads) {
  const int SA = D + 4, d4 = D / 4;
  for (int i = threadIdx.x; i < VQ_MAX_K * d4; i += nthreads) {
    const int code = i / d4, k4 = i - code * d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (code < K) v = __ldg(reinterpret_cast<const float4*>(codebook + static_cast<size_t>(code) * D) + k4);
    *reinterpret_cast<float4*>(es + code * SA + 4 * k4) = v;
  }
  __syncthreads();
  if (threadIdx.x < VQ_MAX_K) {
    const float* e = es + threadIdx.x * SA;
    float s = 0.f;
    for (int k = 0; k < D; k += 16) {
      float4 q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = *reinterpret_cast<const float4*>(e + k + 4 * i);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s = fmaf(q[i].x, q[i].x, s);
        s = fmaf(q[i].y, q[i].y, s);
        s = fmaf(q[i].z, q[i].z, s);
        s = fmaf(q[i].w, q[i].w, s);
      }
    }
    ee[threadIdx.x] = (static_cast<int>(threadIdx.x) < K) ? s : __int_as_float(0x7f800000);
  }
}

// ---- kernel 1: tensor-core prefilter -------------------------------------------------------------------------------
// Approximate distances d'_j = |e_j|^2 - 2 <x, e_j> for all 128 codes of a pass on the tensor cores (single-pass TF32),
// then every code whose d' lies within a guaranteed error margin of the minimum becomes a CANDIDATE: a 128-bit mask per
// vector.  Margin: TF32 truncation gives |dot' - dot| <= 2^-9 sum|x_k e_k| <= 2^-9 |x||e|, so
// |d' + |x|^2 - d| <= 2^-8 |x||e| (+ fp32 roundings, 2^-23-scale); any code j whose exact distance is <= the exact
// distance of the approximate winner satisfies d'_j <= d'_min + 2^-7 |x| max|e|.  The kernel uses 1.25 x that plus
// 2^-18 (|x|^2 + max|e|^2), so the exact argmin -- and every code tied with it -- is always among the candidates.
// A vector with ONE candidate is final when no exact distance is asked for (direct = 1): its index is written here.
// (A first version of this prefilter on mma.sync m16n8k8 TF32 took 77 us at N = 67 840: the legacy tensor path peaks
// near 256 MAC/clk/SM for TF32, a 31 us floor for this problem.)
// D[vector][code] = X^T E^T as tcgen05.mma kind::tf32 with the accumulator in tensor memory:
//   A = the z tile, M = 128 vectors x K = 32 channels per stage.  z is BCHW, i.e. the VECTOR index is the contiguous
//       one, and kind::tf32 takes K-major operands only (an MN-major tf32 descriptor yields zeros: measured with
//       tools/probes/tf32_mn_probe.cu).  So 512 loader threads (thread = vector x channel group) read z with plain
//       4-byte loads (the row pitch of z, HW * 4 bytes, is not a multiple of 16: neither TMA nor wider accesses apply;
//       a warp reads 128 consecutive bytes of a channel row), keep four stages in flight in registers and store each
//       stage as their piece of the K-major, 128-byte-swizzled operand tile (st.shared.v4, conflict free);
//   B = the codebook, N = 128 codes, K-major fp32, resident in shared memory (one TMA load per CTA, 128 KB);
//   D = 128 lanes (vectors) x 128 columns (codes) fp32, double buffered in TMEM.
// Warp roles: warps 0-15 load and transpose, warps 16-19 read the accumulator (thread = vector: all 128 approximate
// distances of a vector sit in one thread, so minimum, margin and candidate mask need no communication), warp 20
// issues TMA and MMA.
constexpr int VQ5_AHEAD = 4;                       // stages in flight in the loaders' registers
constexpr int VQ5_OPS = 4;                         // operand tiles in shared memory
constexpr int VQ5_CH = 32;                         // channels per stage
constexpr int VQ5_STAGE_BYTES = VQ5_CH * VQ_TILE_V * 4;   // 16 KB
constexpr int VQ5_LOAD_WARPS = 16;                 // loader warps: 4 channel groups x 4 vector quarters
constexpr int VQ5_CPT = VQ5_CH / (VQ5_LOAD_WARPS / 4);   // channels per loader thread and stage
constexpr int VQ5_THREADS = 32 * (VQ5_LOAD_WARPS + 5);
constexpr int VQ5_XX_RING = 4;

__global__ void __launch_bounds__(VQ5_THREADS, 1)
vq_prefilter_tc_kernel(const __grid_constant__ CUtensorMap tmCB, const float* __restrict__ z, int B, int D, int HW, int K,
                       uint4* __restrict__ mask_out, long long* __restrict__ idx_out, int code0, int direct) {
  using namespace sm100;
  extern __shared__ uint8_t vq5_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(vq5_raw) + 1023) & ~uintptr_t(1023));
  const int nkb = D / VQ5_CH;                                  // k-blocks of the codebook = stages per tile
  uint8_t* cb = smem;                                          // [nkb][128 rows][128 B]  (K-major, 128-byte swizzle)
  uint8_t* ops = smem + 8 * 16384;                             // [VQ5_OPS][128 rows][128 B] (K-major, 128-byte swizzle)
  float* ee = reinterpret_cast<float*>(ops + VQ5_OPS * VQ5_STAGE_BYTES);   // [128]
  float* xxs = ee + VQ_MAX_K;                                  // [VQ5_XX_RING][channel groups][128] partial |x|^2
  uint64_t* bars = reinterpret_cast<uint64_t*>(xxs + VQ5_XX_RING * (VQ5_LOAD_WARPS / 4) * VQ_TILE_V);
  uint64_t* cb_bar = bars;
  uint64_t* full = bars + 1;                                   // [VQ5_OPS]
  uint64_t* empty = full + VQ5_OPS;                            // [VQ5_OPS]
  uint64_t* tfull = empty + VQ5_OPS;                           // [2]
  uint64_t* tempty = tfull + 2;                                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long N = static_cast<long long>(B) * HW;
  const int num_tiles = static_cast<int>((N + VQ_TILE_V - 1) / VQ_TILE_V);
  const int grid = gridDim.x, bid = blockIdx.x;
  const int my_tiles = (num_tiles - bid + grid - 1) / grid;
  const int total = my_tiles * nkb;                            // stages this CTA streams

  if (t == 0) {
    mbar_init(cb_bar, 1);
    for (int i = 0; i < VQ5_OPS; ++i) {
      mbar_init(full + i, VQ5_LOAD_WARPS);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull + i, 1);
      mbar_init(tempty + i, 128);
    }
    fence_barrier_init();
  }
  if (warp == VQ5_LOAD_WARPS + 4) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < VQ5_LOAD_WARPS) {
    // ================= loaders: thread = (vector v, channel group h): channels h*CPT .. h*CPT+CPT-1 of every stage.
    // VQ5_AHEAD stages are in flight in REGISTERS (plain 4-byte loads: a warp reads 128 consecutive bytes of one channel
    // row; wider accesses were measured slower because the 16-byte phase of a row changes from channel to channel).
    const int v = t & 127, h = t >> 7;
    const long long pitch = static_cast<long long>(HW) * 4;                // bytes between two channels of a vector
    const unsigned int uN = static_cast<unsigned int>(N), uHW = static_cast<unsigned int>(HW);
    int l_j = 0, l_tile = 0, l_c = 0;          // load cursor
    const char* l_ptr = reinterpret_cast<const char*>(z);
    bool l_valid = false;
    auto load = [&](float (&dst)[VQ5_CPT]) {
      if (l_j < total) {
        if (l_c == 0) {
          const unsigned int n = static_cast<unsigned int>(bid + l_tile * grid) * VQ_TILE_V + v;   // N < 2^31 (checked on the host)
          l_valid = n < uN;
          const unsigned int b = l_valid ? n / uHW : 0u;
          const unsigned int pos = l_valid ? n - b * uHW : 0u;
          l_ptr = reinterpret_cast<const char*>(z + (static_cast<long long>(b) * D + h * VQ5_CPT) * HW + pos);
        }
        const char* p = l_ptr;
#pragma unroll
        for (int k = 0; k < VQ5_CPT; ++k) {
          dst[k] = l_valid ? __ldg(reinterpret_cast<const float*>(p)) : 0.f;
          p += pitch;
        }
        l_ptr += VQ5_CH * pitch;
        if (++l_c == nkb) {
          l_c = 0;
          ++l_tile;
        }
      }
      ++l_j;
    };
    float x[VQ5_AHEAD][VQ5_CPT];
#pragma unroll
    for (int u = 0; u < VQ5_AHEAD; ++u) load(x[u]);
    int c_tile = 0, c_c = 0;
    float xx = 0.f;                             // this thread's share of |x|^2 (only the candidate margin uses it)
    const uint32_t my_row = smem_u32(ops) + (v >> 3) * 1024 + (v & 7) * 128;
    uint32_t my_chunk[VQ5_CPT / 4];
#pragma unroll
    for (int c = 0; c < VQ5_CPT / 4; ++c) my_chunk[c] = my_row + (((h * (VQ5_CPT / 4) + c) ^ (v & 7)) << 4);
    for (int j0 = 0; j0 < total; j0 += VQ5_AHEAD) {
#pragma unroll
      for (int u = 0; u < VQ5_AHEAD; ++u) {
        const int j = j0 + u;
        if (j < total) {
#pragma unroll
          for (int k = 0; k < VQ5_CPT; ++k) xx = fmaf(x[u][k], x[u][k], xx);
          const int op = j % VQ5_OPS;
          if (j >= VQ5_OPS) {                   // one polling lane per warp
            if (lane == 0) mbar_wait(empty + op, ((j / VQ5_OPS) - 1) & 1);
            __syncwarp();
          }
#pragma unroll
          for (int c = 0; c < VQ5_CPT / 4; ++c)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my_chunk[c] + op * VQ5_STAGE_BYTES), "f"(x[u][4 * c]),
                         "f"(x[u][4 * c + 1]), "f"(x[u][4 * c + 2]), "f"(x[u][4 * c + 3])
                         : "memory");
          load(x[u]);                           // stage j + VQ5_AHEAD takes over these registers
          if (++c_c == nkb) {
            xxs[((c_tile & (VQ5_XX_RING - 1)) * (VQ5_LOAD_WARPS / 4) + h) * VQ_TILE_V + v] = xx;
            xx = 0.f;
            c_c = 0;
            ++c_tile;
          }
          fence_proxy_async_smem();             // generic-proxy writes -> visible to the tensor core's async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(full + op);   // one arrival per warp
        }
      }
    }
  } else if (warp < VQ5_LOAD_WARPS + 4) {
    // ================= accumulator readers: thread = vector v; columns = codes
    const int q = warp - VQ5_LOAD_WARPS;
    const int v = q * 32 + lane;
    mbar_wait(cb_bar, 0);
    {
      float s = 0.f;                            // |e_v|^2: sequential chain over k, read from the swizzled K-major tile
      const uint8_t* row = cb + (v >> 3) * 1024 + (v & 7) * 128;
      for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 e = *reinterpret_cast<const float4*>(row + kb * 16384 + ((c ^ (v & 7)) << 4));
          s = fmaf(e.x, e.x, s);
          s = fmaf(e.y, e.y, s);
          s = fmaf(e.z, e.z, s);
          s = fmaf(e.w, e.w, s);
        }
      }
      ee[v] = (v < K) ? s : __int_as_float(0x7f800000);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float eemax = 0.f;
    for (int j = 0; j < K; ++j) eemax = fmaxf(eemax, ee[j]);
    const float se = sqrtf(eemax);
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int buf = ti & 1;
      if (lane == 0) mbar_wait(tfull + buf, (ti >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      const float* xp = xxs + (ti & (VQ5_XX_RING - 1)) * (VQ5_LOAD_WARPS / 4) * VQ_TILE_V + v;
      float xxv = 0.f;
#pragma unroll
      for (int w = 0; w < VQ5_LOAD_WARPS / 4; ++w) xxv += xp[w * VQ_TILE_V];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 128;
      float m = __int_as_float(0x7f800000);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fminf(m, fmaf(-2.0f, __uint_as_float(r[j]), ee[ch * 32 + j]));
      }
      const float margin = 1.25f * 0.0078125f * sqrtf(xxv) * se + 3.814697265625e-6f * (xxv + eemax);
      const float thr = m + margin;
      uint32_t mk[4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + ch * 32, r);
        tmem_ld_wait();
        uint32_t bits = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (fmaf(-2.0f, __uint_as_float(r[j]), ee[ch * 32 + j]) <= thr) bits |= 1u << j;
        mk[ch] = bits;
      }
      tc_fence_before();
      mbar_arrive(tempty + buf);                // the MMA warp may overwrite this accumulator
      const long long n = static_cast<long long>(bid + ti * grid) * VQ_TILE_V + v;
      if (n < N) {
        mask_out[n] = make_uint4(mk[0], mk[1], mk[2], mk[3]);
        if (direct && __popc(mk[0]) + __popc(mk[1]) + __popc(mk[2]) + __popc(mk[3]) == 1) {
          const int code = mk[0] ? __ffs(mk[0]) - 1 : (mk[1] ? 31 + __ffs(mk[1]) : (mk[2] ? 63 + __ffs(mk[2]) : 95 + __ffs(mk[3])));
          idx_out[n] = code0 + code;
        }
      }
    }
  } else if (warp == VQ5_LOAD_WARPS + 4 && lane == 0) {
    // ================= TMA (codebook, once) + MMA issue
    prefetch_tensormap(&tmCB);
    mbar_arrive_expect_tx(cb_bar, static_cast<uint32_t>(nkb) * 16384u);
    for (int i = 0; i < nkb; ++i) {
      const int kb = (i + bid) % nkb;   // staggered: the CTAs do not all pull the same L2 lines at the same time
      tma_load_2d(cb + kb * 16384, &tmCB, cb_bar, kb * VQ5_CH, 0, kEvictLast);
    }
    constexpr uint32_t idesc = make_idesc_tf32_f32(128, 128, 0, 0);
    const uint32_t ops_a = smem_u32(ops), cb_a = smem_u32(cb);
    mbar_wait(cb_bar, 0);
    int j = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int buf = ti & 1;
      if (ti >= 2) mbar_wait(tempty + buf, ((ti >> 1) - 1) & 1);
      tc_fence_after();
      for (int c = 0; c < nkb; ++c, ++j) {
        const int op = j % VQ5_OPS;
        mbar_wait(full + op, (j / VQ5_OPS) & 1);
        tc_fence_after();
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          const uint64_t da = make_smem_desc_sw128(ops_a + op * VQ5_STAGE_BYTES + kg * 32);
          const uint64_t db = make_smem_desc_sw128(cb_a + c * 16384 + kg * 32);
          umma_tf32(tmem_base + buf * 128, da, db, idesc, (c | kg) != 0 ? 1u : 0u);
        }
        tc_commit(empty + op);                  // operand tile reusable once these MMAs have read it
      }
      tc_commit(tfull + buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == VQ5_LOAD_WARPS + 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- kernel 2: exact distances of the candidates --------------------------------------------------------------------
// One work item = one (vector, candidate code): the oracle's sequential fmaf chain over the channels
// (d = fl(fl(|x|^2 + |e|^2) - fl(2 <x, e>))), then an atomicMin of (ordered d, code) per vector: smallest distance,
// lowest index among equals = torch.argmin.  With direct = 1 vectors with a single candidate are already final and
// contribute no item.  Items of a 512-vector tile are enumerated through a prefix sum of the candidate counts, so the
// chains of consecutive vectors sit in consecutive lanes (coalesced x reads from L2, where the prefilter just left z).
//
// dynamic smem: es[128][D+4] | ee[128] | best[512] (u64) | pref[513] | cmask[512][4]
__global__ void __launch_bounds__(VQ_EXACT_THREADS, 1)
vq_exact_kernel(const float* __restrict__ z, const float* __restrict__ codebook, int B, int D, int HW, int K,
                const uint4* __restrict__ mask_in, long long* __restrict__ idx_out,
                float* __restrict__ dmin_out, int code0, int merge, int direct) {
  codebook += static_cast<size_t>(code0) * D;
  extern __shared__ __align__(16) float vq_smem[];
  const int SA = D + 4;
  float* es = vq_smem;
  float* ee = es + static_cast<size_t>(VQ_MAX_K) * SA;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(ee + VQ_MAX_K);    // [512]
  int* pref = reinterpret_cast<int*>(best + VQ_EXACT_TILE);                             // [513] (+3 pad)
  uint32_t* cmask = reinterpret_cast<uint32_t*>(pref + VQ_EXACT_TILE + 4);              // [512][4]
  __shared__ int s_warp_tot[VQ_EXACT_THREADS / 32];

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long N = static_cast<long long>(B) * HW;
  const int num_tiles = static_cast<int>((N + VQ_EXACT_TILE - 1) / VQ_EXACT_TILE);
  stage_codebook(codebook, D, K, es, ee, VQ_EXACT_THREADS);
  __syncthreads();

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long n0 = static_cast<long long>(tile) * VQ_EXACT_TILE;
    // ---- candidate counts and their exclusive prefix sum
    int cnt = 0;
    if (t < VQ_EXACT_TILE) {
      uint4 mk = make_uint4(0u, 0u, 0u, 0u);
      if (n0 + t < N) mk = mask_in[n0 + t];
      cnt = __popc(mk.x) + __popc(mk.y) + __popc(mk.z) + __popc(mk.w);
      if (direct && cnt == 1) cnt = 0;
      *reinterpret_cast<uint4*>(cmask + 4 * t) = mk;
      best[t] = ~0ull;
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = lane < VQ_EXACT_THREADS / 32 ? s_warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      if (lane < VQ_EXACT_THREADS / 32) s_warp_tot[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int warp_base = warp > 0 ? s_warp_tot[warp - 1] : 0;
    if (t < VQ_EXACT_TILE) pref[t] = warp_base + inc - cnt;
    const int total = s_warp_tot[VQ_EXACT_TILE / 32 - 1];
    if (t == 0) pref[VQ_EXACT_TILE] = total;
    __syncthreads();

    // ---- exact chains
    for (int w = t; w < total; w += VQ_EXACT_THREADS) {
      int lo = 0, hi = VQ_EXACT_TILE;          // largest v with pref[v] <= w (and a non-empty item range)
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pref[mid] <= w) lo = mid; else hi = mid;
      }
      const int v = lo;
      int r = w - pref[v];
      int code = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t word = cmask[4 * v + q];
        const int c = __popc(word);
        if (r >= 0 && r < c) code = 32 * q + static_cast<int>(__fns(word, 0, r + 1));
        r -= c;                                   // negative once found
      }
      const long long n = n0 + v;
      const long long b = n / HW;
      const int pos = static_cast<int>(n - b * HW);
      const float* xp = z + (b * D) * HW + pos;
      const float* ep = es + code * SA;
      float dot = 0.f, xx = 0.f;      // two sequential chains over the channel index, as the oracle evaluates them
#pragma unroll 1
      for (int k = 0; k < D; k += 64) {
        float xv[64];                 // 64 independent loads in flight per thread: the chain is latency bound on L2
#pragma unroll
        for (int i = 0; i < 64; ++i) xv[i] = __ldg(xp + static_cast<long long>(k + i) * HW);
#pragma unroll
        for (int i4 = 0; i4 < 16; ++i4) {
          const float4 e = *reinterpret_cast<const float4*>(ep + k + 4 * i4);
          dot = fmaf(xv[4 * i4 + 0], e.x, dot);
          xx = fmaf(xv[4 * i4 + 0], xv[4 * i4 + 0], xx);
          dot = fmaf(xv[4 * i4 + 1], e.y, dot);
          xx = fmaf(xv[4 * i4 + 1], xv[4 * i4 + 1], xx);
          dot = fmaf(xv[4 * i4 + 2], e.z, dot);
          xx = fmaf(xv[4 * i4 + 2], xv[4 * i4 + 2], xx);
          dot = fmaf(xv[4 * i4 + 3], e.w, dot);
          xx = fmaf(xv[4 * i4 + 3], xv[4 * i4 + 3], xx);
        }
      }
      const float d = __fsub_rn(__fadd_rn(xx, ee[code]), __fmul_rn(2.0f, dot));
      atomicMin(&best[v], (static_cast<unsigned long long>(f32_ord(d)) << 32) | static_cast<unsigned int>(code));
    }
    __syncthreads();
    if (t < VQ_EXACT_TILE && cnt > 0) {
      const long long n = n0 + t;
      const unsigned long long key = best[t];
      const float d = f32_unord(static_cast<uint32_t>(key >> 32));
      const bool keep_old = merge && !(d < dmin_out[n]);
      if (!keep_old) {
        idx_out[n] = code0 + static_cast<int>(key & 0xffffffffull);
        if (dmin_out) dmin_out[n] = d;
      }
    }
    __syncthreads();   // best / pref / cmask are rewritten by the next tile
  }
}

// quantized / straight-through value / one-hot / loss partial sums / code histogram.
// One thread per (b, pos) vector-channel element, indexed in BCHW order so both the read
// of z and the write of quantized are coalesced.
__global__ void vq_finish_kernel(const float* __restrict__ z, const float* __restrict__ codebook,
                                 const long long* __restrict__ idx, int B, int D, int HW, int K,
                                 float* __restrict__ quantized, float* __restrict__ encodings,
                                 double* __restrict__ sq_sum, unsigned int* __restrict__ counts) {
  const long long total = static_cast<long long>(B) * D * HW;
  const long long N = static_cast<long long>(B) * HW;
  double local = 0.0;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pos = static_cast<int>(e % HW);
    const long long bc = e / HW;
    const int c = static_cast<int>(bc % D);
    const long long b = bc / D;
    const long long n = b * HW + pos;
    const int code = static_cast<int>(idx[n]);
    const float q = __ldg(codebook + static_cast<size_t>(code) * D + c);
    const float x = z[e];
    const float diff = __fsub_rn(q, x);
    if (quantized) quantized[e] = __fadd_rn(x, diff);  // inputs + (quantized - inputs).detach()   (:49)
    local += static_cast<double>(diff) * static_cast<double>(diff);
    if (c == 0) {
      if (counts) atomicAdd(&counts[code], 1u);
    }
  }
  // one-hot encodings (N, K): written by a second grid-stride loop over N*K
  if (encodings) {
    const long long tot2 = N * K;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < tot2;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long n = e / K;
      const int k = static_cast<int>(e - n * K);
      encodings[e] = (static_cast<int>(idx[n]) == k) ? 1.0f : 0.0f;
    }
  }
  // block reduction of the squared error
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && sq_sum) atomicAdd(sq_sum, v);
  }
}

// loss = q_latent_loss + commitment_cost * e_latent_loss, both = mse(quantized, inputs)  (:43-45)
// perplexity = exp(-sum p log(p + 1e-10)), p = mean(encodings, 0)                        (:50-51)
__global__ void vq_scalars_kernel(const double* __restrict__ sq_sum, const unsigned int* __restrict__ counts,
                                  long long numel, long long N, int K, float commitment_cost,
                                  float* __restrict__ loss_out, float* __restrict__ perplexity_out) {
  if (threadIdx.x == 0 && loss_out) {
    const float mse = static_cast<float>(*sq_sum / static_cast<double>(numel));
    *loss_out = __fadd_rn(mse, __fmul_rn(commitment_cost, mse));
  }
  if (perplexity_out) {
    float s = 0.f;
    for (int k = threadIdx.x; k < K; k += 32) {
      const float p = static_cast<float>(counts[k]) / static_cast<float>(N);
      s += p * logf(p + 1e-10f);
    }
    s = warp_sum(s);
    if (threadIdx.x == 0) *perplexity_out = expf(-s);
  }
}

// get_codebook_entry: pure gather.  hw > 0: output is BCHW (B = n_vec / hw); hw == 0: (N, D) rows.
__global__ void vq_gather_kernel(const long long* __restrict__ idx, const float* __restrict__ codebook, long long n_vec,
                                 int D, int HW, int K, float* __restrict__ out, int* __restrict__ bad_index) {
  const long long total = n_vec * D;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long n;
    int c;
    if (HW > 0) {
      const int pos = static_cast<int>(e % HW);
      const long long bc = e / HW;
      c = static_cast<int>(bc % D);
      n = (bc / D) * HW + pos;
    } else {
      n = e / D;
      c = static_cast<int>(e - n * D);
    }
    const long long code = idx[n];
    if (code < 0 || code >= K) {
      if (bad_index) atomicExch(bad_index, 1);
      out[e] = 0.f;
    } else {
      out[e] = __ldg(codebook + static_cast<size_t>(code) * D + c);
    }
  }
}

}  // namespace

// stream-ordered scratch pool of the current device: keeps its memory across synchronisations (release threshold =
// max), so a call costs two pointer bumps instead of a driver allocation
static int scratch_pool(cudaMemPool_t* out) {
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  MGV_CHECK_CUDA(cudaGetDevice(&dev));
  MGV_REQUIRE(dev >= 0 && dev < 64, "vq: device ordinal %d out of range", dev);
  if (!pools[dev]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t p = nullptr;
    MGV_CHECK_CUDA(cudaMemPoolCreate(&p, &props));
    unsigned long long keep = ~0ull;
    MGV_CHECK_CUDA(cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep));
    pools[dev] = p;
  }
  *out = pools[dev];
  return MGV_OK;
}

int vq_argmin(const float* z, const float* codebook, int B, int D, int HW, int K, long long* idx_out, float* dmin_out,
              cudaStream_t stream) {
  MGV_REQUIRE(B >= 0 && HW > 0 && static_cast<long long>(B) * HW < (1LL << 31) - 1024, "vq_argmin: B=%d HW=%d", B, HW);
  if (B == 0) return MGV_OK;   // empty batch: nothing to do (pointers may be null)
  MGV_REQUIRE(z && codebook && idx_out, "vq_argmin: null pointer");
  MGV_REQUIRE(K >= 1 && K <= 65536, "vq_argmin: num_embeddings=%d unsupported (1..65536)", K);
  MGV_REQUIRE(D >= 64 && D % 64 == 0 && D <= 256, "vq_argmin: embedding_dim=%d must be 64, 128, 192 or 256", D);
  // codebooks beyond one 128-code pass: the passes hand the running minimum to each other through dmin_out
  // (caller-owned, so that concurrent streams / devices never share scratch memory)
  MGV_REQUIRE(K <= VQ_MAX_K || dmin_out != nullptr, "vq_argmin: num_embeddings=%d > %d needs a dmin buffer of B*HW floats", K, VQ_MAX_K);
  const long long N = static_cast<long long>(B) * HW;
  const int n_sm = num_sms();
  const int tiles1 = static_cast<int>((N + VQ_TILE_V - 1) / VQ_TILE_V);
  const int tiles2 = static_cast<int>((N + VQ_EXACT_TILE - 1) / VQ_EXACT_TILE);
  const int grid1 = tiles1 < n_sm ? tiles1 : n_sm;
  const int grid2 = tiles2 < n_sm ? tiles2 : n_sm;
  const size_t cb_bytes = (static_cast<size_t>(VQ_MAX_K) * (D + 4) + VQ_MAX_K) * sizeof(float);
  const size_t smem1 = 8 * 16384 + VQ5_OPS * VQ5_STAGE_BYTES + (VQ_MAX_K + VQ5_XX_RING * (VQ5_LOAD_WARPS / 4) * VQ_TILE_V) * sizeof(float) + 256 + 1024;
  const size_t smem2 = cb_bytes + VQ_EXACT_TILE * 8 + (VQ_EXACT_TILE + 4) * 4 + VQ_EXACT_TILE * 16;
  static unsigned long long attr_mask = 0;   // per device
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(vq_prefilter_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem1)));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(vq_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  // candidate masks (16 B per vector): stream-ordered scratch, so concurrent streams and devices never share it
  void* scratch = nullptr;
  cudaMemPool_t pool = nullptr;
  MGV_TRY(scratch_pool(&pool));
  MGV_CHECK_CUDA(cudaMallocFromPoolAsync(&scratch, static_cast<size_t>(N) * 16, pool, stream));
  uint4* masks = static_cast<uint4*>(scratch);
  // a vector with a single candidate is final unless its exact distance is needed (dmin output, multi-pass merge)
  const int direct = (dmin_out == nullptr && K <= VQ_MAX_K) ? 1 : 0;
  for (int code0 = 0; code0 < K; code0 += VQ_MAX_K) {
    const int kc = (K - code0 < VQ_MAX_K) ? K - code0 : VQ_MAX_K;
    CUtensorMap tmCB;
    MGV_TRY(make_tmap_2d_f32(&tmCB, codebook + static_cast<size_t>(code0) * D, D, kc, static_cast<uint64_t>(D) * 4, VQ5_CH, VQ_MAX_K));
    vq_prefilter_tc_kernel<<<grid1, VQ5_THREADS, smem1, stream>>>(tmCB, z, B, D, HW, kc, masks, idx_out, code0, direct);
    vq_exact_kernel<<<grid2, VQ_EXACT_THREADS, smem2, stream>>>(z, codebook, B, D, HW, kc, masks, idx_out, dmin_out, code0,
                                                                code0 > 0 ? 1 : 0, direct);
  }
  const cudaError_t launch_err = cudaGetLastError();
  MGV_CHECK_CUDA(cudaFreeAsync(scratch, stream));
  MGV_CHECK_CUDA(launch_err);
  return MGV_OK;
}

int vq_finish(const float* z, const float* codebook, const long long* idx, int B, int D, int HW, int K,
              float commitment_cost, float* quantized, float* encodings, float* loss_out, float* perplexity_out,
              void* workspace /* >= 8 + 4*K bytes, 8-byte aligned */, cudaStream_t stream) {
  if (B == 0) return MGV_OK;
  MGV_REQUIRE(z && codebook && idx && workspace, "vq_finish: null pointer");
  double* sq = static_cast<double*>(workspace);
  unsigned int* counts = reinterpret_cast<unsigned int*>(sq + 1);
  MGV_CHECK_CUDA(cudaMemsetAsync(workspace, 0, 8 + 4 * static_cast<size_t>(K), stream));
  const long long total = static_cast<long long>(B) * D * HW;
  int blocks = static_cast<int>((total + 255) / 256);
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  vq_finish_kernel<<<blocks, 256, 0, stream>>>(z, codebook, idx, B, D, HW, K, quantized, encodings, sq, counts);
  MGV_CHECK_CUDA(cudaGetLastError());
  vq_scalars_kernel<<<1, 32, 0, stream>>>(sq, counts, total, static_cast<long long>(B) * HW, K, commitment_cost, loss_out,
                                           perplexity_out);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

int vq_gather(const long long* idx, const float* codebook, long long n_vec, int D, int HW, int K, float* out,
              int* bad_index_flag, cudaStream_t stream) {
  MGV_REQUIRE(HW >= 0 && (HW == 0 || n_vec % HW == 0), "vq_gather: n_vec=%lld not a multiple of HW=%d", n_vec, HW);
  if (n_vec == 0) return MGV_OK;
  MGV_REQUIRE(idx && codebook && out, "vq_gather: null pointer");
  const long long total = n_vec * D;
  int blocks = static_cast<int>((total + 255) / 256);
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  vq_gather_kernel<<<blocks, 256, 0, stream>>>(idx, codebook, n_vec, D, HW, K, out, bad_index_flag);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

}  // namespace mgv
