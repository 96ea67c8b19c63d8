// tcgen05 GEMM / implicit-GEMM convolution kernel (see gemm_tc.cuh).
//
// CTA = 192 threads: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA
// issuer (one lane), warps 2..5 = epilogue (each owns one 32-lane TMEM quarter).
// Tile = 128 (M) x BN (N) x 64 (K) per pipeline stage; the ring of `stages` stages is
// filled by TMA with the 128-byte swizzle and drained by tcgen05.mma reading smem
// descriptors; the fp32 accumulator lives in TMEM (BN columns) and is read back with
// tcgen05.ld for the fused epilogue.
//
// Weight streaming (decode step): when launched with programmatic dependent launch the
// producer issues the B (weight) loads of the first ring pass BEFORE griddepcontrol.wait,
// so weights stream from HBM while the upstream kernel is still finishing; only the A
// (activation) loads wait for the dependency.
#include <stdlib.h>
#include "gemm_tc.cuh"

namespace mgv {

using namespace sm100;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 192;

struct KParams {
  int M, N, K;
  int kb_per_split;  // 64-wide k blocks handled by one blockIdx.z
  int stages;
  int epi;
  const float* bias;
  void* out;
  const void* resid;
  long long ldo;
  // conv
  int a_mode;
  int H, W, Cin, wb, hb, tiles_x, tiles_y, stride, pad;
  int pad_y, taps_x;              // pad = x padding; filter taps per row (K = taps_y * taps_x * Cin)
  int oscale, oy, ox;             // output pixel (y, x) is stored at (y * oscale + oy, x * oscale + ox) of an (H*oscale) x (W*oscale) image
  int gn_tile_base, gn_tiles_img; // GroupNorm slot of tile (img, ty, tx) = img * gn_tiles_img + gn_tile_base + ty * tiles_x + tx
  float* gn_sum;
  int gn_group_ch;
  int evict_first_w;   // L2 evict-first hint on the weight operand
  int transpose_out;   // swap-AB: weights are the A operand, output written transposed
};

// ---- per-warp 32 x 64-byte transposer (2 KB of shared memory per warp)
// The accumulator layout gives every lane one output ROW, so a direct store instruction touches 32 rows x 16 bytes:
// 32 half-filled sectors per instruction.  Measured on the 80x848 convolutions: with the epilogue's global stores
// and residual loads removed the kernel ran 26 % faster, with its arithmetic removed as well only 2 % more -- the
// memory instructions were the cost.  Staging through shared memory turns every instruction into 8 rows x 64
// contiguous bytes (16 full sectors).  "own" = lane l holds the 64 bytes of row l; "spread" = instruction i of lane l
// holds 16-byte piece (l & 3) of row 8*i + (l >> 2).  The XOR keeps both access patterns bank-conflict free.
constexpr int EPI_STAGE_BYTES = 32 * 64;
// explicit shared-space accesses (through a generic pointer parameter the compiler emits generic ST.E / LD.E)
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t stage_own(uint32_t buf, int lane, int j) {
  return buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
}
__device__ __forceinline__ uint32_t stage_spread(uint32_t buf, int lane, int i) {
  const int r = 8 * i + (lane >> 2), c = lane & 3;
  return buf + r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
}
__device__ __forceinline__ void own_to_spread(uint8_t* stage, int lane, const uint4 (&in)[4], uint4 (&out)[4]) {
  const uint32_t buf = smem_u32(stage);
#pragma unroll
  for (int j = 0; j < 4; ++j) sts_v4(stage_own(buf, lane, j), in[j]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = lds_v4(stage_spread(buf, lane, i));
  __syncwarp();
}
__device__ __forceinline__ void spread_to_own(uint8_t* stage, int lane, const uint4 (&in)[4], uint4 (&out)[4]) {
  const uint32_t buf = smem_u32(stage);
#pragma unroll
  for (int i = 0; i < 4; ++i) sts_v4(stage_spread(buf, lane, i), in[i]);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) out[j] = lds_v4(stage_own(buf, lane, j));
  __syncwarp();
}

// Row-major epilogue of one accumulator tile: this thread owns accumulator row (TMEM lane) `quarter*32+lane`
// and walks the BN columns in chunks of 32 (bias / GELU / residual / bf16 or fp32 store / split-K reduction /
// GroupNorm partial sums into this warp's shared-memory bins).  `stage` = this warp's EPI_STAGE_BYTES transposer
// buffer (nullptr: direct, uncoalesced path).
template <int BN>
__device__ __forceinline__ void epilogue_rows(const KParams& p, uint32_t tmem_base, int quarter, int lane, bool row_ok,
                                              long long out_row, int n0, bool add_bias, float* s_bins,
                                              int c_begin = 0, int c_end = BN / 32, uint8_t* stage = nullptr) {
    // element offsets / validity of the 4 rows this lane touches in the spread pattern
    long long srow[4];
    bool sok[4];
    if (stage != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int src = 8 * i + (lane >> 2);
        srow[i] = __shfl_sync(0xffffffffu, out_row, src);
        sok[i] = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, src) != 0;
      }
    }
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= p.N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (add_bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(b4 + j);
          v[4 * j + 0] += b.x;
          v[4 * j + 1] += b.y;
          v[4 * j + 2] += b.z;
          v[4 * j + 3] += b.w;
        }
      }
      if (p.epi == EPI_BF16 || p.epi == EPI_BF16_GELU || p.epi == EPI_BF16_RESID) {
        if (p.epi == EPI_BF16_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (p.epi == EPI_BF16_RESID && stage != nullptr) {
          uint4 sp[4], own[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            sp[i] = make_uint4(0u, 0u, 0u, 0u);
            if (sok[i])
              sp[i] = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.resid) + srow[i] + col0) + (lane & 3));
          }
          spread_to_own(stage, lane, sp, own);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {own[j].x, own[j].y, own[j].z, own[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16x2(w[e]);
              v[8 * j + 2 * e] += f.x;
              v[8 * j + 2 * e + 1] += f.y;
            }
          }
        } else if (p.epi == EPI_BF16_RESID && row_ok) {
          const uint4* r4 = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.resid) + out_row + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = __ldg(r4 + j);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16x2(w[e]);
              v[8 * j + 2 * e] += f.x;
              v[8 * j + 2 * e + 1] += f.y;
            }
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        if (stage != nullptr) {
          uint4 own[4], sp[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) own[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          own_to_spread(stage, lane, own, sp);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (sok[i]) *(reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + srow[i] + col0) + (lane & 3)) = sp[i];
        } else if (row_ok) {
          uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + out_row + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) o4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        if (p.gn_sum != nullptr) {
          // GroupNorm statistics of the stored (bf16-rounded) values; a tile never spans images.
          // Per lane: (sum, sumsq) of each channel group of this 32-column chunk -> 2*(32/gch) <= 16 values;
          // a transposing butterfly (16 shuffles) leaves one fully reduced value per lane pair, which goes to
          // the per-CTA shared-memory bins (flushed once per CTA after the chunk loop).
          const int gch = p.gn_group_ch;          // 4, 8 or 16 channels per group
          float st[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) st[i] = 0.f;
          if (row_ok) {
            float sq[16], sm[16];                 // per channel pair
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 f = unpack_bf16x2(pk[j]);
              sm[j] = f.x + f.y;
              sq[j] = f.x * f.x + f.y * f.y;
            }
            if (gch == 4) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                st[2 * g] = sm[2 * g] + sm[2 * g + 1];
                st[2 * g + 1] = sq[2 * g] + sq[2 * g + 1];
              }
            } else if (gch == 8) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                st[2 * g] = (sm[4 * g] + sm[4 * g + 1]) + (sm[4 * g + 2] + sm[4 * g + 3]);
                st[2 * g + 1] = (sq[4 * g] + sq[4 * g + 1]) + (sq[4 * g + 2] + sq[4 * g + 3]);
              }
            } else {
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  a += sm[8 * g + e];
                  b += sq[8 * g + e];
                }
                st[2 * g] = a;
                st[2 * g + 1] = b;
              }
            }
          }
          // butterfly: after the steps with offsets 16, 8, 4, 2 lane l holds value index (l >> 1) & 15
#pragma unroll
          for (int step = 0; step < 4; ++step) {
            const int off = 16 >> step, n = 8 >> step;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              const float send = upper ? st[i] : st[i + n];
              const float keep = upper ? st[i + n] : st[i];
              st[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          st[0] += __shfl_xor_sync(0xffffffffu, st[0], 1);
          const int vidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
          const int nvals = 2 * (32 / gch);
          // each warp owns its bins (one writer per bin): plain, deterministic accumulation
          if ((lane & 1) == 0 && vidx < nvals) s_bins[quarter * 64 + c * nvals + vidx] += st[0];
        }
      } else if (stage != nullptr && p.epi != EPI_F32_ATOMIC) {
        // fp32 rows are 128 bytes per chunk: two passes of 16 columns (64 bytes) through the transposer
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 own[4], sp[4];
          if (p.epi == EPI_F32_RESID) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              sp[i] = make_uint4(0u, 0u, 0u, 0u);
              if (sok[i])
                sp[i] = *(reinterpret_cast<const uint4*>(static_cast<const float*>(p.resid) + srow[i] + col0 + 16 * h) + (lane & 3));
            }
            spread_to_own(stage, lane, sp, own);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[16 * h + 4 * j + 0] += __uint_as_float(own[j].x);
              v[16 * h + 4 * j + 1] += __uint_as_float(own[j].y);
              v[16 * h + 4 * j + 2] += __uint_as_float(own[j].z);
              v[16 * h + 4 * j + 3] += __uint_as_float(own[j].w);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            own[j] = make_uint4(__float_as_uint(v[16 * h + 4 * j]), __float_as_uint(v[16 * h + 4 * j + 1]),
                                __float_as_uint(v[16 * h + 4 * j + 2]), __float_as_uint(v[16 * h + 4 * j + 3]));
          own_to_spread(stage, lane, own, sp);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (sok[i]) *(reinterpret_cast<uint4*>(static_cast<float*>(p.out) + srow[i] + col0 + 16 * h) + (lane & 3)) = sp[i];
        }
      } else if (row_ok) {
        float* o = static_cast<float*>(p.out) + out_row + col0;
        if (p.epi == EPI_F32_ATOMIC) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            atomicAdd(reinterpret_cast<float4*>(o) + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        } else {
          if (p.epi == EPI_F32_RESID) {
            const float4* r4 = reinterpret_cast<const float4*>(static_cast<const float*>(p.resid) + out_row + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 q = r4[j];
              v[4 * j + 0] += q.x;
              v[4 * j + 1] += q.y;
              v[4 * j + 2] += q.z;
              v[4 * j + 3] += q.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
    }
}

// DECODE = true: the swap-AB weight-streaming form only (plain A operand, transposed epilogue) -- the decode
// loop launches this kernel ~100 times per position for a few microseconds each, so its code is kept minimal.
template <int BN, bool DECODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bins = reinterpret_cast<float*>(tmem_slot + 4);   // [4 epilogue warps][64] GroupNorm partial sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int total_kb = p.K / BK;
  int nkb = total_kb - kb0;
  if (nkb > p.kb_per_split) nkb = p.kb_per_split;

  int m0 = 0, img = 0, x0 = 0, y0 = 0;
  long long stats_slot = blockIdx.x;
  if (DECODE || p.a_mode == A_PLAIN) {
    m0 = blockIdx.x * BM;
  } else {
    int t = blockIdx.x;
    int tx = t % p.tiles_x;
    t /= p.tiles_x;
    int ty = t % p.tiles_y;
    img = t / p.tiles_y;
    x0 = tx * p.wb;
    y0 = ty * p.hb;
    stats_slot = static_cast<long long>(img) * p.gn_tiles_img + p.gn_tile_base + ty * p.tiles_x + tx;
  }

  if (!DECODE && threadIdx.x >= 64) {   // 128 epilogue threads clear their warp's 64 bins
    s_bins[threadIdx.x - 64] = 0.f;
    s_bins[threadIdx.x + 64] = 0.f;
  }
  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // let the dependent grid start its prologue (and prefetch its weights) now; its griddepcontrol.wait
  // still blocks until this grid has completed and flushed
  pdl_launch_dependents();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const uint64_t hint_w = p.evict_first_w ? kEvictFirst : kEvictNormal;
      const bool swap_ab = DECODE || p.transpose_out;
      const uint64_t hint_b = swap_ab ? kEvictNormal : hint_w;
      const uint64_t hint_a = swap_ab ? hint_w : kEvictNormal;
      const int cblocks = (DECODE || p.a_mode == A_PLAIN) ? 1 : (p.Cin / BK);
      auto load_a = [&](int kb, int s, uint64_t* bar = nullptr) {
        uint8_t* dst = smem + s * STAGE_BYTES;
        if (bar == nullptr) bar = &full_bar[s];
        if (DECODE || p.a_mode == A_PLAIN) {
          tma_load_2d(dst, &tmA, bar, (kb0 + kb) * BK, m0, hint_a);
        } else {
          const int kk = kb0 + kb;
          const int tap = kk / cblocks;
          const int c0 = (kk - tap * cblocks) * BK;
          const int dy = tap / p.taps_x, dx = tap - dy * p.taps_x;
          // input coordinate of the tile's first pixel for this tap
          const int xi = x0 * p.stride + dx - p.pad;
          const int yi = y0 * p.stride + dy - p.pad_y;
          tma_load_4d(dst, &tmA, bar, c0, xi, yi, img, hint_a);
        }
      };
      auto load_b = [&](int kb, int s, uint64_t* bar = nullptr) {
        uint8_t* dst = smem + s * STAGE_BYTES + A_STAGE_BYTES;
        if (bar == nullptr) bar = &full_bar[s];
        tma_load_2d(dst, &tmB, bar, (kb0 + kb) * BK, n0, hint_b);
      };
      if (DECODE && nkb <= stages) {
        // Decode fast path: the whole K slice is resident, so there is no ring to manage -- one barrier for the weight
        // tiles (issued before the grid dependency resolves), one for the activation tiles, one MMA burst, one commit.
        // (Per k-block hand-offs cost the issuing threads ~0.25 us each: see gemm_tc_persist_kernel.)
        uint64_t* w_bar = &full_bar[0];
        uint64_t* x_bar = &empty_bar[0];
        mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(nkb) * A_STAGE_BYTES);
        for (int kb = 0; kb < nkb; ++kb) load_a(kb, kb, w_bar);
        pdl_wait();
        mbar_arrive_expect_tx(x_bar, static_cast<uint32_t>(nkb) * B_STAGE_BYTES);
        for (int kb = 0; kb < nkb; ++kb) load_b(kb, kb, x_bar);
      } else {
      const int pre = nkb < stages ? nkb : stages;
      for (int kb = 0; kb < pre; ++kb) {  // weights first: they do not depend on the upstream grid
        mbar_arrive_expect_tx(&full_bar[kb], STAGE_BYTES);
        if (swap_ab) load_a(kb, kb); else load_b(kb, kb);
      }
      pdl_wait();
      for (int kb = 0; kb < pre; ++kb) {
        if (swap_ab) load_b(kb, kb); else load_a(kb, kb);
      }
      for (int kb = pre; kb < nkb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        load_a(kb, s);
        load_b(kb, s);
      }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BN);
      if (DECODE && nkb <= stages) {
        mbar_wait(&full_bar[0], 0);    // weights
        mbar_wait(&empty_bar[0], 0);   // activations
        tc_fence_after();
        const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem));
        const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem) + A_STAGE_BYTES);
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t off = static_cast<uint64_t>((kb * STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, da0 + off + 2 * k, db0 + off + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
      } else
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row
          umma_bf16(tmem_base, make_smem_desc_sw128(a_addr + k * 32), make_smem_desc_sw128(b_addr + k * 32),
                    idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      tc_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // =========================== epilogue ===========================
    pdl_wait();  // residual / output buffers belong to the upstream grid
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    const int r_local = quarter * 32 + lane;
    bool row_ok;
    long long out_row;  // element offset of this thread's output row
    if (DECODE || p.a_mode == A_PLAIN) {
      const int row = m0 + r_local;
      row_ok = row < p.M;
      out_row = static_cast<long long>(row) * p.ldo;
    } else {
      const int ly = r_local / p.wb, lx = r_local - ly * p.wb;
      const int x = x0 + lx, y = y0 + ly;
      row_ok = (x < p.W) && (y < p.H);
      out_row = ((static_cast<long long>(img) * p.H * p.oscale + y * p.oscale + p.oy) * (p.W * p.oscale) + x * p.oscale + p.ox) * p.ldo;
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const bool add_bias = p.bias != nullptr && (p.epi != EPI_F32_ATOMIC || blockIdx.z == 0);
    if (DECODE || p.transpose_out) {
      // swap-AB: this thread owns output feature `feat`; accumulator columns are batch rows.  For a fixed
      // batch row the 32 lanes of a warp touch 32 consecutive features: coalesced stores / reductions.
      const int feat = m0 + r_local;
      const bool feat_ok = feat < p.M;
      const float bval = (add_bias && feat_ok) ? __ldg(p.bias + feat) : 0.f;
      // NOTE: this kernel runs for a few microseconds per launch in the decode loop, so the code executed
      // here is kept COMPACT (instruction fetch of a bloated, fully unrolled epilogue costs more than the math).
      if (p.epi == EPI_F32_ATOMIC) {
        float* outf = static_cast<float*>(p.out) + feat;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (col0 + j < p.N && feat_ok)
              atomicAdd(outf + static_cast<long long>(col0 + j) * p.ldo, __uint_as_float(r[j]) + bval);
          }
        }
      } else {
        const int ncols = (p.N - n0 < BN) ? p.N - n0 : BN;
#pragma unroll 1
        for (int j = 0; j < ncols; ++j) {   // one accumulator column (= batch row) at a time
          const uint32_t raw = tmem_ld_32x32_x1(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + j);
          tmem_ld_wait();
          if (feat_ok) {
            const float v = __uint_as_float(raw) + bval;
            const long long o = static_cast<long long>(n0 + j) * p.ldo + feat;
            switch (p.epi) {
              case EPI_F32: static_cast<float*>(p.out)[o] = v; break;
              case EPI_F32_RESID: static_cast<float*>(p.out)[o] = static_cast<const float*>(p.resid)[o] + v; break;
              case EPI_BF16_GELU: static_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16(gelu_erf(v)); break;
              case EPI_BF16_RESID:
                static_cast<__nv_bfloat16*>(p.out)[o] =
                    __float2bfloat16(v + __bfloat162float(static_cast<const __nv_bfloat16*>(p.resid)[o]));
                break;
              default: static_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16(v); break;
            }
          }
        }
      }
    } else if constexpr (!DECODE) {
      epilogue_rows<BN>(p, tmem_base, quarter, lane, row_ok, out_row, n0, add_bias, s_bins);
    }  // !DECODE
    if (!DECODE && p.gn_sum != nullptr) {
      // one plain store per CTA and bin into this tile's slot (deterministic; folded by vqvae_gn_finalize)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int nbins = (BN / p.gn_group_ch) * 2;
      const int b = threadIdx.x - 64;
      const int groups_total = p.N / p.gn_group_ch;
      const int g0 = n0 / p.gn_group_ch;
      if (b >= 0 && b < nbins && g0 + (b >> 1) < groups_total)
        p.gn_sum[(stats_slot * groups_total + g0 + (b >> 1)) * 2 + (b & 1)] =
            (s_bins[b] + s_bins[64 + b]) + (s_bins[128 + b] + s_bins[192 + b]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}


// ---------------------------------------------------------------- persistent variant (large grids)
// One CTA per SM loops over output tiles.  The smem ring keeps streaming across tile boundaries and the
// accumulator is double buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the MMAs of
// tile i+1 -- the canonical warp-specialised Blackwell GEMM.  Used for the implicit-GEMM convolutions and the
// prefill GEMMs; the decode GEMMs (one tile per CTA) use gemm_tc_kernel<BN, true>.
struct PTile {
  int m0, n0, img, x0, y0;
  long long stats_slot;   // index of this tile's GroupNorm partial-sum slot
};

__device__ __forceinline__ PTile decode_tile(const KParams& p, int tile, int n_tiles, int BN) {
  PTile t;
  const int mt = tile / n_tiles;
  t.n0 = (tile - mt * n_tiles) * BN;
  t.m0 = 0; t.img = 0; t.x0 = 0; t.y0 = 0;
  t.stats_slot = mt;
  if (p.a_mode == A_PLAIN) {
    t.m0 = mt * BM;
  } else {
    int r = mt;
    const int tx = r % p.tiles_x;
    r /= p.tiles_x;
    const int ty = r % p.tiles_y;
    t.img = r / p.tiles_y;
    t.x0 = tx * p.wb;
    t.y0 = ty * p.hb;
    t.stats_slot = static_cast<long long>(t.img) * p.gn_tiles_img + p.gn_tile_base + ty * p.tiles_x + tx;
  }
  return t;
}

constexpr int PERSIST_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int KSUB = 2;                // k-blocks per pipeline stage (one barrier hand-off per KSUB k-blocks)

// Measured on the 80x848 128->128 convolutions: with the loads, the MMAs and the epilogue all removed the
// one-k-block-per-stage version of this kernel still took 56 % of its time -- the single MMA-issuing thread needs
// ~500 cycles per k-block for its barrier wait, descriptor arithmetic (runtime modulo / division by the stage
// count), four issues and the commit, while the tensor core finishes the four 128x128x16 MMAs in 256.  Hence: two
// k-blocks per stage (half the hand-offs), stage / phase / tap counters instead of divisions, descriptors advanced
// by adding to a precomputed base.
// MT = M tiles per CTA.  MT = 2 (BN <= 128): the CTA owns two 128-row tiles that share every B tile, i.e. a 256 x BN
// output block -- (256 + BN) operand rows per k-block instead of 2 x (128 + BN).  Once the issue loop was fixed the
// 128 x 128 convolutions became bound by the L2 -> SM operand feed (~53 B/cycle/SM), so fewer bytes per FLOP is time.
template <int BN, int MT>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const KParams p, int m_tiles, int n_tiles) {
  static_assert(MT == 1 || (MT == 2 && BN <= 128), "two M tiles need 4 x BN <= 512 TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int KB_BYTES = MT * A_STAGE_BYTES + B_STAGE_BYTES;     // one k-block: MT A tiles, then the B tile
  constexpr int STAGE_BYTES = KSUB * KB_BYTES;
  constexpr int ACC_COLS = MT * BN;                                // TMEM columns of one accumulator buffer
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bins = reinterpret_cast<float*>(tmem_slot + 4);   // [MT][4 quarters][64] GroupNorm partial sums
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_bins + MT * 256);   // [8 epilogue warps][EPI_STAGE_BYTES], 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_groups = (m_tiles + MT - 1) / MT;
  const int total_tiles = m_groups * n_tiles;        // CTA-level tiles (MT x 128 rows each)
  const int nkb = p.K / BK;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], PERSIST_THREADS - 64);   // every epilogue thread arrives
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * ACC_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // M tile `mt` (0..MT-1) of CTA-level tile `tile`; tiles past the end decode to out-of-range coordinates (TMA zero
  // fill, stores masked by the caller)
  auto sub_tile = [&](int tile, int mt) {
    const int mg = tile / n_tiles;
    const int nt = tile - mg * n_tiles;
    return decode_tile(p, (mg * MT + mt) * n_tiles + nt, n_tiles, BN);
  };

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const int cblocks = (p.a_mode == A_PLAIN) ? 1 : (p.Cin / BK);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        PTile t[MT];
        int xb[MT], yb[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          t[mt] = sub_tile(tile, mt);
          xb[mt] = t[mt].x0 * p.stride - p.pad;
          yb[mt] = t[mt].y0 * p.stride - p.pad_y;
        }
        int cb = 0, dx = 0, dy = 0;          // conv: channel block and filter tap of the next k-block
        for (int kb = 0; kb < nkb; kb += KSUB) {
          const int nsub = (nkb - kb < KSUB) ? nkb - kb : KSUB;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(nsub) * KB_BYTES);
          uint8_t* dst = smem + s * STAGE_BYTES;
          for (int j = 0; j < nsub; ++j, dst += KB_BYTES) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (p.a_mode == A_PLAIN)
                tma_load_2d(dst + mt * A_STAGE_BYTES, &tmA, &full_bar[s], (kb + j) * BK, t[mt].m0, kEvictNormal);
              else
                tma_load_4d(dst + mt * A_STAGE_BYTES, &tmA, &full_bar[s], cb * BK, xb[mt] + dx, yb[mt] + dy, t[mt].img,
                            kEvictNormal);
            }
            if (p.a_mode != A_PLAIN && ++cb == cblocks) {
              cb = 0;
              if (++dx == p.taps_x) {
                dx = 0;
                ++dy;
              }
            }
            tma_load_2d(dst + MT * A_STAGE_BYTES, &tmB, &full_bar[s], (kb + j) * BK, t[0].n0, kEvictLast);
          }
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BN);
      const uint64_t desc_a0 = make_smem_desc_sw128(smem_u32(smem));
      const uint64_t desc_b0 = make_smem_desc_sw128(smem_u32(smem) + MT * A_STAGE_BYTES);
      int s = 0, lt = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int as = lt & 1;
        mbar_wait(&tmem_empty_bar[as], ((lt >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * ACC_COLS;
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; kb += KSUB) {
          const int nsub = (nkb - kb < KSUB) ? nkb - kb : KSUB;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          // descriptor start-address field counts 16-byte units
          uint64_t da = desc_a0 + static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
          uint64_t db = desc_b0 + static_cast<uint64_t>((s * STAGE_BYTES) >> 4);
          for (int j = 0; j < nsub; ++j, da += KB_BYTES >> 4, db += KB_BYTES >> 4) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {   // 16 bf16 = 32 bytes along K inside the 128-byte swizzle row
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                umma_bf16(tmem_acc + mt * BN, da + mt * (A_STAGE_BYTES >> 4) + 2 * k, db + 2 * k, idesc, acc);
              acc = 1;
            }
          }
          tc_commit(&empty_bar[s]);
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
        tc_commit(&tmem_full_bar[as]);
      }
    }
  } else {
    // =========================== epilogue ===========================
    // warps 2..9: TMEM lane quarter = warp & 3 (hardware rule).  MT = 1: the two warps of a quarter split the
    // columns; MT = 2: each takes one of the two M tiles (all columns).
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int NCH = BN / 32;
    const int mt_mine = (MT == 2) ? half : 0;
    const int c_begin = (MT == 2) ? 0 : ((NCH >= 2) ? half * (NCH / 2) : 0);
    const int c_end = (MT == 2) ? NCH : ((NCH >= 2) ? c_begin + NCH / 2 : (half == 0 ? 1 : 0));
    const int r_local = quarter * 32 + lane;
    const int et = threadIdx.x - 64;      // epilogue thread index 0..255
    float* my_bins = s_bins + mt_mine * 256;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const PTile t = sub_tile(tile, mt_mine);
      const bool tile_ok = (tile / n_tiles) * MT + mt_mine < m_tiles;
      const int as = lt & 1;
      bool row_ok;
      long long out_row;
      if (p.a_mode == A_PLAIN) {
        const int row = t.m0 + r_local;
        row_ok = tile_ok && row < p.M;
        out_row = static_cast<long long>(row) * p.ldo;
      } else {
        const int ly = r_local / p.wb, lx = r_local - ly * p.wb;
        const int x = t.x0 + lx, y = t.y0 + ly;
        row_ok = tile_ok && (x < p.W) && (y < p.H);
        out_row = ((static_cast<long long>(t.img) * p.H * p.oscale + y * p.oscale + p.oy) * (p.W * p.oscale) + x * p.oscale + p.ox) * p.ldo;
      }
      if (p.gn_sum != nullptr) {   // [MT][4 quarters][64] bins; cleared after the previous flush barrier
        s_bins[et] = 0.f;
        if (MT == 2) s_bins[256 + et] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(&tmem_full_bar[as], (lt >> 1) & 1);
      tc_fence_after();
      epilogue_rows<BN>(p, tmem_base + as * ACC_COLS + mt_mine * BN, quarter, lane, row_ok, out_row, t.n0, p.bias != nullptr,
                        my_bins, c_begin, c_end, s_stage + (warp - 2) * EPI_STAGE_BYTES);
      // accumulator fully read into registers: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[as]);
      if (p.gn_sum != nullptr) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int nbins = (BN / p.gn_group_ch) * 2;
        const int groups_total = p.N / p.gn_group_ch;
        // thread et < 128 flushes M tile 0, 128 <= et < 256 flushes M tile 1 (MT = 2)
        const int fm = (MT == 2) ? (et >> 7) : 0;
        const int eb = (MT == 2) ? (et & 127) : et;
        if (fm < MT && eb < nbins) {
          const PTile ft = sub_tile(tile, fm);
          const int g0 = ft.n0 / p.gn_group_ch;
          const bool f_ok = (tile / n_tiles) * MT + fm < m_tiles;
          const float* bsrc = s_bins + fm * 256;
          if (f_ok && g0 + (eb >> 1) < groups_total)
            p.gn_sum[(ft.stats_slot * groups_total + g0 + (eb >> 1)) * 2 + (eb & 1)] =
                (bsrc[eb] + bsrc[64 + eb]) + (bsrc[128 + eb] + bsrc[192 + eb]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // bins may be cleared for the next tile
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * ACC_COLS);
  }
}

// ---------------------------------------------------------------- SIMT reference
__global__ void gemm_ref_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                KParams p, long long lda, int n_img, int Hin, int Win) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(p.M) * p.N) return;
  const int n = static_cast<int>(idx % p.N);
  const long long m = idx / p.N;
  float acc = 0.f;
  if (p.a_mode == A_PLAIN) {
    for (int k = 0; k < p.K; ++k)
      acc = fmaf(__bfloat162float(A[m * lda + k]), __bfloat162float(B[static_cast<long long>(n) * p.K + k]), acc);
  } else {
    const int x = static_cast<int>(m % p.W);
    const int y = static_cast<int>((m / p.W) % p.H);
    const int img = static_cast<int>(m / (static_cast<long long>(p.W) * p.H));
    const int ntaps = p.K / p.Cin;
    for (int tap = 0; tap < ntaps; ++tap) {
      const int dy = tap / p.taps_x, dx = tap % p.taps_x;
      const int xi = x * p.stride + dx - p.pad, yi = y * p.stride + dy - p.pad_y;
      if (xi < 0 || yi < 0 || xi >= Win || yi >= Hin) continue;
      const __nv_bfloat16* a = A + ((static_cast<long long>(img) * Hin + yi) * Win + xi) * p.Cin;
      const __nv_bfloat16* b = B + static_cast<long long>(n) * p.K + tap * p.Cin;
      for (int c = 0; c < p.Cin; ++c) acc = fmaf(__bfloat162float(a[c]), __bfloat162float(b[c]), acc);
    }
  }
  if (p.bias) acc += p.transpose_out ? p.bias[m] : p.bias[n];
  long long o = p.transpose_out ? static_cast<long long>(n) * p.ldo + m : m * p.ldo + n;
  if (p.a_mode != A_PLAIN) {
    const int x = static_cast<int>(m % p.W), y = static_cast<int>((m / p.W) % p.H);
    const long long img = m / (static_cast<long long>(p.W) * p.H);
    o = ((img * p.H * p.oscale + y * p.oscale + p.oy) * (p.W * p.oscale) + x * p.oscale + p.ox) * p.ldo + n;
  }
  switch (p.epi) {
    case EPI_BF16: static_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16(acc); break;
    case EPI_BF16_GELU: static_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16(gelu_erf(acc)); break;
    case EPI_BF16_RESID:
      static_cast<__nv_bfloat16*>(p.out)[o] =
          __float2bfloat16(acc + __bfloat162float(static_cast<const __nv_bfloat16*>(p.resid)[o]));
      break;
    case EPI_F32: static_cast<float*>(p.out)[o] = acc; break;
    case EPI_F32_RESID: static_cast<float*>(p.out)[o] = static_cast<const float*>(p.resid)[o] + acc; break;
    case EPI_F32_ATOMIC: static_cast<float*>(p.out)[o] += acc; break;
  }
}

int fill_params(const GemmArgs& a, KParams& p) {
  MGV_REQUIRE(a.A && a.B && a.out, "gemm: null operand");
  MGV_REQUIRE(a.K > 0 && a.K % BK == 0, "gemm: K=%d must be a positive multiple of %d", a.K, BK);
  MGV_REQUIRE(a.N > 0 && (a.transpose_out || a.N % 32 == 0), "gemm: N=%d must be a multiple of 32", a.N);
  MGV_REQUIRE(!a.transpose_out || (a.a_mode == A_PLAIN && a.gn_sum == nullptr), "gemm: transpose_out is a plain-GEMM mode");
  MGV_REQUIRE(a.M > 0, "gemm: M=%d", a.M);
  MGV_REQUIRE(a.split_k >= 1, "gemm: split_k=%d", a.split_k);
  MGV_REQUIRE(a.split_k == 1 || a.epi == EPI_F32_ATOMIC, "gemm: split_k>1 needs EPI_F32_ATOMIC");
  if (a.epi == EPI_F32_RESID || a.epi == EPI_BF16_RESID) MGV_REQUIRE(a.resid, "gemm: residual epilogue without resid");
  memset(&p, 0, sizeof(p));
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.epi = a.epi;
  p.bias = a.bias;
  p.out = a.out;
  p.resid = a.resid;
  p.ldo = a.ldo ? a.ldo : (a.transpose_out ? a.M : a.N);
  p.a_mode = a.a_mode;
  p.evict_first_w = a.weights_evict_first ? 1 : 0;
  p.transpose_out = a.transpose_out ? 1 : 0;

  p.gn_sum = a.gn_sum;
  p.gn_group_ch = a.gn_group_ch;
  if (a.gn_sum) {
    MGV_REQUIRE(a.a_mode == A_CONV3x3 || a.n_img > 0, "gemm: gn_sum needs image geometry");
    MGV_REQUIRE(a.gn_group_ch == 4 || a.gn_group_ch == 8 || a.gn_group_ch == 16, "gemm: gn_group_ch=%d", a.gn_group_ch);
    MGV_REQUIRE(a.epi == EPI_BF16 || a.epi == EPI_BF16_RESID, "gemm: gn_sum needs a bf16 epilogue");
  }
  if (a.a_mode == A_CONV3x3) {
    MGV_REQUIRE(a.n_img > 0 && a.H > 0 && a.W > 0 && a.Cin > 0 && a.Cin % BK == 0, "conv: bad geometry");
    MGV_REQUIRE(a.taps_x >= 1 && a.taps_x <= 3 && a.K % (a.taps_x * a.Cin) == 0, "conv: K=%d is not taps_y * %d * Cin", a.K, a.taps_x);
    MGV_REQUIRE(a.out_scale >= 1 && a.out_oy >= 0 && a.out_oy < a.out_scale && a.out_ox >= 0 && a.out_ox < a.out_scale, "conv: output placement");
    MGV_REQUIRE(a.out_scale == 1 || a.resid == nullptr, "conv: strided output placement has no residual form");
    MGV_REQUIRE(a.M == a.n_img * a.H * a.W, "conv: M mismatch");
    MGV_REQUIRE(a.stride == 1 || a.stride == 2, "conv: stride");
    p.H = a.H;
    p.W = a.W;
    p.Cin = a.Cin;
    p.stride = a.stride;
    p.pad = a.pad;
    p.pad_y = a.pad_y >= 0 ? a.pad_y : a.pad;
    p.taps_x = a.taps_x;
    p.oscale = a.out_scale;
    p.oy = a.out_oy;
    p.ox = a.out_ox;
    p.wb = conv_tile_width(a.H, a.W);
    p.hb = BM / p.wb;
    p.tiles_x = ceil_div(a.W, p.wb);
    p.tiles_y = ceil_div(a.H, p.hb);
    p.gn_tiles_img = a.gn_tiles_img > 0 ? a.gn_tiles_img : p.tiles_x * p.tiles_y;
    p.gn_tile_base = a.gn_tile_base;
  }
  return MGV_OK;
}

template <int BN, bool DECODE>
int launch_tc(const GemmArgs& a, KParams& p) {
  constexpr int STAGE_BYTES = A_STAGE_BYTES + BN * BK * 2;
  const int total_kb = a.K / BK;
  p.kb_per_split = ceil_div(total_kb, a.split_k);
  const int splits = ceil_div(total_kb, p.kb_per_split);
  int max_stages = (200 * 1024) / STAGE_BYTES;
  if (a.max_stages > 0 && a.max_stages < max_stages) max_stages = a.max_stages;
  p.stages = p.kb_per_split < max_stages ? p.kb_per_split : max_stages;
  if (p.stages < 1) p.stages = 1;
  const size_t smem = static_cast<size_t>(p.stages) * STAGE_BYTES + (2 * p.stages + 1) * 8 + 16 + 1024 + 1024;

  CUtensorMap tmA, tmB;
  if (a.a_mode == A_PLAIN) {
    const int64_t lda = a.lda ? a.lda : a.K;
    MGV_TRY(make_tmap_2d_bf16(&tmA, a.A, a.K, a.M, lda * 2, BK, BM));
  } else {
    MGV_TRY(make_tmap_nhwc_bf16(&tmA, a.A, a.Cin, a.Win, a.Hin, a.n_img, BK, p.wb, p.hb, a.stride));
  }
  MGV_TRY(make_tmap_2d_bf16(&tmB, a.B, a.K, a.N, static_cast<uint64_t>(a.K) * 2, BK, BN));

  static unsigned long long attr_mask = 0;   // per device (and per template instantiation)
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, DECODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // ask for the full shared-memory carve-out so that several CTAs (small rings) can share an SM
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, DECODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
  }
  dim3 grid;
  if (a.a_mode == A_PLAIN)
    grid = dim3(ceil_div(a.M, BM), ceil_div(a.N, BN), splits);
  else
    grid = dim3(p.tiles_x * p.tiles_y * a.n_img, ceil_div(a.N, BN), 1);
  LaunchCfg lc(grid, dim3(GEMM_THREADS), smem, a.stream, a.pdl);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, gemm_tc_kernel<BN, DECODE>, tmA, tmB, p));
  return MGV_OK;
}


template <int BN, int MT>
int launch_persist(const GemmArgs& a, KParams& p) {
  constexpr int STAGE_BYTES = KSUB * (MT * A_STAGE_BYTES + BN * BK * 2);
  p.kb_per_split = a.K / BK;
  int stages = (200 * 1024) / STAGE_BYTES;   // one CTA per SM: the ring takes what the epilogue staging leaves
  if (stages > 6) stages = 6;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * STAGE_BYTES + (2 * stages + 4) * 8 + 16 + MT * 1024 + 8 * EPI_STAGE_BYTES + 1024;
  CUtensorMap tmA, tmB;
  int m_tiles;
  if (a.a_mode == A_PLAIN) {
    const int64_t lda = a.lda ? a.lda : a.K;
    MGV_TRY(make_tmap_2d_bf16(&tmA, a.A, a.K, a.M, lda * 2, BK, BM));
    m_tiles = ceil_div(a.M, BM);
  } else {
    MGV_TRY(make_tmap_nhwc_bf16(&tmA, a.A, a.Cin, a.Win, a.Hin, a.n_img, BK, p.wb, p.hb, a.stride));
    m_tiles = p.tiles_x * p.tiles_y * a.n_img;
  }
  MGV_TRY(make_tmap_2d_bf16(&tmB, a.B, a.K, a.N, static_cast<uint64_t>(a.K) * 2, BK, BN));
  const int n_tiles = ceil_div(a.N, BN);
  static unsigned long long attr_mask = 0;   // per device (and per template instantiation)
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_persist_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_persist_kernel<BN, MT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
  }
  const int total = ceil_div(m_tiles, MT) * n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  LaunchCfg lc(dim3(grid), dim3(PERSIST_THREADS), smem, a.stream, false);
  MGV_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, gemm_tc_persist_kernel<BN, MT>, tmA, tmB, p, m_tiles, n_tiles));
  return MGV_OK;
}

}  // namespace

int gemm_bf16_tc(const GemmArgs& a) {
  KParams p;
  MGV_TRY(fill_params(a, p));
  if (a.a_mode == A_CONV3x3) MGV_REQUIRE(a.split_k == 1, "conv: split_k unsupported");
  // large plain / conv problems: persistent kernel with double-buffered TMEM accumulators
  static const bool no_persist = getenv("MGV_NO_PERSIST") != nullptr;
  if (!a.transpose_out && a.split_k == 1 && !no_persist) {
    const long long tiles = a.a_mode == A_PLAIN ? static_cast<long long>(ceil_div(a.M, BM)) * ceil_div(a.N, a.bn)
                                                : static_cast<long long>(p.tiles_x) * p.tiles_y * a.n_img * ceil_div(a.N, a.bn);
    static const bool no_mt2 = getenv("MGV_NO_MT2") != nullptr;
    if (tiles > 2LL * num_sms()) {
      // two M tiles per CTA (256 x BN output block) whenever the tile count allows it
      const bool mt2 = !no_mt2 && tiles > 4LL * num_sms();
      switch (a.bn) {
        case 32: return launch_persist<32, 1>(a, p);
        case 64: return mt2 ? launch_persist<64, 2>(a, p) : launch_persist<64, 1>(a, p);
        case 128: return mt2 ? launch_persist<128, 2>(a, p) : launch_persist<128, 1>(a, p);
        case 256: return launch_persist<256, 1>(a, p);
        default: break;
      }
    }
  }
  if (a.transpose_out) {   // decode form: compact kernel
    switch (a.bn) {
      case 32: return launch_tc<32, true>(a, p);
      case 64: return launch_tc<64, true>(a, p);
      case 128: return launch_tc<128, true>(a, p);
      case 256: return launch_tc<256, true>(a, p);
      default: set_error("gemm: bn=%d not in {32,64,128,256}", a.bn); return MGV_ERR_INVALID;
    }
  }
  switch (a.bn) {
    case 32: return launch_tc<32, false>(a, p);
    case 64: return launch_tc<64, false>(a, p);
    case 128: return launch_tc<128, false>(a, p);
    case 256: return launch_tc<256, false>(a, p);
    default: set_error("gemm: bn=%d not in {32,64,128,256}", a.bn); return MGV_ERR_INVALID;
  }
}

int gemm_bf16_ref(const GemmArgs& a) {
  KParams p;
  MGV_TRY(fill_params(a, p));
  const long long total = static_cast<long long>(a.M) * a.N;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  gemm_ref_kernel<<<static_cast<unsigned>(blocks), threads, 0, a.stream>>>(
      static_cast<const __nv_bfloat16*>(a.A), static_cast<const __nv_bfloat16*>(a.B), p, a.lda ? a.lda : a.K, a.n_img,
      a.Hin, a.Win);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

}  // namespace mgv
