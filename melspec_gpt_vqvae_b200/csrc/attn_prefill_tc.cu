// Causal self-attention of the prefill / teacher-forced pass on the 5th-generation tensor cores
// (reference: CausalSelfAttention.forward transformer/minGPT.py:70-92 -- att = softmax(mask(q k^T / sqrt(d))), y = att v).
//
// One CTA = one (sequence, head).  T <= 272 rows are cut into at most three 128-row query tiles FROM THE END
// (T = 265: rows 0..8, 9..136, 137..264), so the short tile is the one that sees the fewest keys and the padded MMA work
// is 16 + 144 + 272 key columns instead of 128 + 256 + 272.  Everything a head needs is resident at once:
//   shared memory  Q tiles | K (all keys) | V (all keys), loaded by TMA straight out of the [B*T, 3C] qkv tensor through
//                  ONE 3-D tensor map (OOB rows of a sequence are zero-filled, 128-byte swizzle);
//   TMEM           S_i = Q_i K^T for all tiles side by side (16 + 144 + 272 fp32 columns) + 64 columns for O_0;
//                  O_i (i >= 1) reuses the first 64 columns of S_i once the softmax has consumed it.
// S:  tcgen05.mma kind::f16, A = Q tile (K-major), B = K rows (K-major), N split into chunks of <= 256 keys.
// softmax: 8 warps, two per TMEM lane quarter (a thread = one query row x every other 32-column chunk; partial row maxima
//          and sums are exchanged through shared memory); exact two-pass form (row maximum, then exp2 / sum) over the
//          causal prefix of the row only (chunks beyond a warp's last key are written as zeros, never loaded; chunks
//          entirely below the diagonal skip the mask arithmetic);
//          P goes to shared memory as bf16 in the K-major 128-byte-swizzled layout of an MMA A operand -- into the
//          Q | K region, which is dead once every S MMA has completed.
// O = P V: A = P (K-major), B = V in the layout TMA delivered it: rows = keys, 64 contiguous head dims = an MN-major
//          operand (descriptor form verified by tools/probes/bf16_mn_probe.cu: SBO = 8 keys x 128 B, K step = 2048 B).
// Order: the LAST tile (the long one) first -- its softmax overlaps the issue of the other tiles' MMAs (one thread issues
// every tcgen05.mma at ~120 cycles apiece) and the kernel's tail is the short tile.
// No [T, T] tensor reaches HBM.  The kernel covers the plain causal mask without the attention-map output and without
// the KV-cache fill; gpt_attention_prefill keeps the mma.sync kernel for those cases.
#include <stdlib.h>
#include "gpt_kernels.cuh"
#include "mgv_sm100.cuh"

namespace mgv {

using namespace sm100;

namespace {

constexpr int AT_THREADS = 320;     // warp 0: TMA + MMA issue, warp 1: TMEM allocation, warps 2..9: softmax / epilogue
constexpr int AT_TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int AT_MAX_TILES = 3;

struct AttnTcParams {
  int T, nh, C;
  int n_tiles;
  int row0[AT_MAX_TILES], nrows[AT_MAX_TILES], kpad[AT_MAX_TILES], scol[AT_MAX_TILES], ocol[AT_MAX_TILES];
  int kmax;           // key rows of K / V in shared memory (= kpad of the last tile)
  int q0_small;       // tile 0 has at most 16 rows: one 16-row box instead of a 128-row one
  int qoff[AT_MAX_TILES], koff;   // byte offsets of the Q tiles and of K in shared memory
  int pk_bytes;       // bytes of the Q | K region (>= the P tiles of the last tile, which reuse it)
  int p1_bytes;       // second P region (tiles at odd distance from the last one), so that consecutive tiles never share one
  int tmem_cols;      // power of two
  __nv_bfloat16* y;   // [B*T, C]
  long long* trace;   // optional (diagnostics): clock64 stamps of CTA 0, [16] per recording thread
  int heads;          // persistent form: B * nh
  int buf_bytes;      // persistent form: bytes of one head's Q | K | V buffer
};

// kind::f16, A = bf16 K-major, B = bf16 K-major (b_mn = 0) or MN-major (b_mn = 1), D = f32
__host__ __device__ constexpr uint32_t at_idesc(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ float max3(float a, float b, float c) {   // one FMNMX3 on sm_100
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm16,
                       const __grid_constant__ CUtensorMap tm32, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // Q tiles, last tile first (a short tile 0 takes 2 KB) } reused for P
  uint8_t* sK = smem + p.koff;                          // [kmax] x 128 B                                       } after the S MMAs
  uint8_t* sV = smem + p.pk_bytes;                      // [kmax] x 128 B: row = key, 128 B = 64 head dims
  uint8_t* sP = smem;
  uint8_t* sP1 = sV + p.kmax * 128;
  float* s_m = reinterpret_cast<float*>(sP1 + p.p1_bytes);    // [2][128] partial row maxima of the two column halves
  float* s_l = s_m + 256;                                     // [2][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_l + 256);
  uint64_t* bar_qk = bars;                              // Q of the last tile + K landed
  uint64_t* bar_q = bars + 1;                           // the other Q tiles landed
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;                           // [0] S of the last tile complete, [1] every S complete
  uint64_t* bar_p = bars + 5;                           // [3] P_i written
  uint64_t* bar_o = bars + 8;                           // [3] O_i complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.nh, h = blockIdx.x - b * p.nh;
  // diagnostics: thread 0 (TMA / MMA issuer) -> slots 0..15, thread 96 (warp 3: quarter 3, half 0) -> 16..31,
  // thread 224 (warp 7: quarter 3, half 1) -> 32..47
  const bool tr = p.trace != nullptr && blockIdx.x == gridDim.x / 2 &&
                  (threadIdx.x == 0 || threadIdx.x == 96 || threadIdx.x == 224);
  long long* trp = p.trace + (threadIdx.x == 0 ? 0 : (threadIdx.x == 96 ? 16 : 32));
  int tri = 0;
#define AT_STAMP() do { if (tr && tri < 16) trp[tri++] = clock64(); } while (0)
  AT_STAMP();

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm);
    prefetch_tensormap(&tm16);
    prefetch_tensormap(&tm32);
    mbar_init(bar_qk, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_v, 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    for (int i = 0; i < AT_MAX_TILES; ++i) {
      mbar_init(&bar_p[i], AT_THREADS - 64);
      mbar_init(&bar_o[i], 1);
    }
    fence_barrier_init();
    // The loads are issued before the TMEM allocation is awaited (shapes small enough for two CTAs per SM: the second
    // one sits in tcgen05.alloc until TMEM is free, with its operands already landing).
    // ---- loads.  The tiles are processed LAST TILE FIRST (it is the long one: its softmax then overlaps the issue of
    // the remaining MMAs, and the tail of the kernel is the short tile), so its Q rows and K go first, on their own barrier.
    // 128-row boxes, then 16-row boxes for the remainder (T = 265: 272 key rows = 2 x 128 + 16; Q tile 0 = 16 rows)
    const int kfull = p.kmax / 128, krem = (p.kmax - kfull * 128) / 16;
    const int last = p.n_tiles - 1;
    const uint32_t q0_bytes = p.q0_small ? 16 * 128 : AT_TILE_BYTES;
    mbar_arrive_expect_tx(bar_qk, (last == 0 ? q0_bytes : AT_TILE_BYTES) + static_cast<uint32_t>(p.kmax) * 128);
    if (last > 0) {
      // the last tile's 32-row quarters go to shared memory in REVERSE order (see the softmax warps: load balance)
      for (int q = 0; q < 4; ++q)
        tma_load_4d(sQ + p.qoff[last] + q * 32 * 128, &tm32, bar_qk, h * GPT_HEAD_DIM, p.row0[last] + (3 - q) * 32, b, 0,
                    kEvictFirst);
    } else {
      tma_load_4d(sQ + p.qoff[last], p.q0_small ? &tm16 : &tm, bar_qk, h * GPT_HEAD_DIM, p.row0[last], b, 0, kEvictFirst);
    }
    for (int j = 0; j < kfull; ++j)
      tma_load_4d(sK + j * AT_TILE_BYTES, &tm, bar_qk, p.C + h * GPT_HEAD_DIM, j * 128, b, 0, kEvictFirst);
    for (int j = 0; j < krem; ++j)
      tma_load_4d(sK + (kfull * 128 + j * 16) * 128, &tm16, bar_qk, p.C + h * GPT_HEAD_DIM, kfull * 128 + j * 16, b, 0,
                  kEvictFirst);
    if (last > 0) {
      mbar_arrive_expect_tx(bar_q, q0_bytes + static_cast<uint32_t>(last - 1) * AT_TILE_BYTES);
      for (int i = last - 1; i >= 0; --i)
        tma_load_4d(sQ + p.qoff[i], (i == 0 && p.q0_small) ? &tm16 : &tm, bar_q, h * GPT_HEAD_DIM, p.row0[i], b, 0,
                    kEvictFirst);
    }
    mbar_arrive_expect_tx(bar_v, static_cast<uint32_t>(p.kmax) * 128);
    for (int j = 0; j < kfull; ++j)
      tma_load_4d(sV + j * AT_TILE_BYTES, &tm, bar_v, 2 * p.C + h * GPT_HEAD_DIM, j * 128, b, 0, kEvictFirst);
    for (int j = 0; j < krem; ++j)
      tma_load_4d(sV + (kfull * 128 + j * 16) * 128, &tm16, bar_v, 2 * p.C + h * GPT_HEAD_DIM, kfull * 128 + j * 16, b, 0,
                  kEvictFirst);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  AT_STAMP();

  if (warp == 0) {
    if (lane == 0) {
      const int last = p.n_tiles - 1;
      AT_STAMP();

      // ---- S_i = Q_i K^T, last tile first (N in chunks of <= 256 keys, 4 K steps of 16 head dims).  Rows of a short
      // tile 0 beyond its 16-row box are whatever shared memory held: they only reach accumulator rows that no one
      // reads unmasked.
      auto issue_s = [&](int i) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(sQ + p.qoff[i]));
        for (int n0 = 0; n0 < p.kpad[i]; n0 += 256) {
          const int n = (p.kpad[i] - n0 < 256) ? p.kpad[i] - n0 : 256;
          const uint64_t db = make_smem_desc_sw128(smem_u32(sK) + n0 * 128);
          const uint32_t idesc = at_idesc(128, n, 0);
#pragma unroll
          for (int k = 0; k < GPT_HEAD_DIM / 16; ++k)
            umma_bf16(tmem_base + p.scol[i] + n0, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
        }
      };
      mbar_wait(bar_qk, 0);
      AT_STAMP();
      tc_fence_after();
      issue_s(last);
      tc_commit(&bar_s[0]);
      AT_STAMP();
      if (last > 0) {
        mbar_wait(bar_q, 0);
        tc_fence_after();
        for (int i = last - 1; i >= 0; --i) issue_s(i);
      }
      tc_commit(&bar_s[1]);
      AT_STAMP();

      // ---- O_i = P_i V as soon as the softmax warps have written P_i
      mbar_wait(bar_v, 0);
      const uint32_t idesc_pv = at_idesc(128, GPT_HEAD_DIM, 1);
      for (int i = last; i >= 0; --i) {
        mbar_wait(&bar_p[i], 0);
        AT_STAMP();
        tc_fence_after();
        const int ksteps = p.kpad[i] / 16;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = make_smem_desc_sw128(smem_u32(((last - i) & 1) ? sP1 : sP) + (ks >> 2) * AT_TILE_BYTES + (ks & 3) * 32);
          const uint64_t db = make_smem_desc_mn_sw128(smem_u32(sV) + ks * 2048, 8192, 1024);
          umma_bf16(tmem_base + p.ocol[i], da, db, idesc_pv, ks != 0 ? 1u : 0u);
        }
        tc_commit(&bar_o[i]);
        AT_STAMP();
      }
    }
  } else if (warp >= 2) {
    // ---- softmax + epilogue: TMEM lane quarter = warp & 3; the two warps of a quarter take alternate 32-column chunks
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rl = quarter * 32 + lane;
    const int last = p.n_tiles - 1;
    // Tile row (TMEM lane) -> sequence row.  A warp's work grows with the number of keys its rows can see, and the lane
    // quarter of a warp (= its SM sub-partition, with its one MUFU unit) is fixed.  With every tile in natural order
    // quarter 3 would get the longest rows of every tile (T = 265: 15 chunks of 32 columns against 9 for quarter 0);
    // the last tile is therefore stored with its quarters reversed: 12 / 11 / 11 / 11.
    auto tile_row = [&](int i) { return (i == last && last > 0) ? (3 - quarter) * 32 + lane : rl; };
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // 1/sqrt(d) (minGPT.py:81) folded with log2(e): probabilities are exp2(s' - max')
    const float scale2 = 1.4426950408889634f * 0.125f;
    mbar_wait(&bar_s[0], 0);
    tc_fence_after();
    AT_STAMP();

    float lsum[AT_MAX_TILES] = {0.f, 0.f, 0.f};
    // normalise and store this thread's half (32 head dims) of output row rl of tile i
    auto epilogue = [&](int i, float l) {
      mbar_wait(&bar_o[i], 0);
      tc_fence_after();
      const int tr_ = tile_row(i);
      const bool ok = tr_ < p.nrows[i];
      const float inv = ok ? 1.0f / l : 0.f;
      __nv_bfloat16* dst = p.y + (static_cast<long long>(b) * p.T + p.row0[i] + tr_) * p.C + h * GPT_HEAD_DIM + half * 32;
      uint32_t r[32];
      tmem_ld_32x32(lane_base + p.ocol[i] + half * 32, r);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
          q.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
          q.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
          q.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + j * 8) = q;
        }
      }
    };

    for (int i = last; i >= 0; --i) {
      if (i == p.n_tiles - 2) {        // the remaining S tiles (issued while the last tile's pass 1 ran)
        mbar_wait(&bar_s[1], 0);
        tc_fence_after();
      }
      const int trow = tile_row(i);
      const int q0row = trow - lane;                        // first tile row of this warp
      const bool ok = trow < p.nrows[i];
      const int row = p.row0[i] + trow;                     // keys 0..row are visible (tril mask, minGPT.py:65-68, :82)
      // last key any row of this warp can see -> number of 32-column chunks to load (warp-uniform)
      const int wlast_rl = (q0row + 31 < p.nrows[i]) ? q0row + 31 : p.nrows[i] - 1;
      const int wkeys = (wlast_rl >= q0row) ? p.row0[i] + wlast_rl + 1 : 0;
      const int nch_load = (wkeys + 31) / 32;
      const int nch_all = (p.kpad[i] + 31) / 32;
      // chunks entirely at or below the diagonal for EVERY row of this warp need no mask (all 32 rows valid)
      const int nch_full = (q0row + 31 < p.nrows[i]) ? (p.row0[i] + q0row + 1) / 32 : 0;
      const uint32_t s_addr = lane_base + p.scol[i];

      // pass 1: exact row maximum of the scores (this thread's chunks, two TMEM loads in flight), combined with the
      // other half's through shared memory
      float m = -INFINITY;
      auto chunk_max = [&](const uint32_t (&r)[32], int c) {
        if (c < nch_full) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) m = max3(m, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j <= row) m = fmaxf(m, __uint_as_float(r[j]));
        }
      };
      for (int c = half; c < nch_load; c += 4) {
        uint32_t r0[32], r1[32];
        const bool two = c + 2 < nch_load;
        tmem_ld_32x32(s_addr + c * 32, r0);
        if (two) tmem_ld_32x32(s_addr + (c + 2) * 32, r1);
        tmem_ld_wait();
        chunk_max(r0, c);
        if (two) chunk_max(r1, c + 2);
      }
      s_m[half * 128 + rl] = m;
      AT_STAMP();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      m = fmaxf(s_m[rl], s_m[128 + rl]) * scale2;
      // P_i goes to the region tile i + 2 used (consecutive tiles alternate between two regions, so P V of tile i + 1 never
      // has to be awaited here): that tile's MMAs finished long ago -- its epilogue runs now.  The last tile (processed
      // first) writes over Q | K, so every S MMA must be done.
      if (i == last) mbar_wait(&bar_s[1], 0);
      else if (i + 2 <= last) epilogue(i + 2, lsum[i + 2]);
      AT_STAMP();

      // pass 2: p = exp2(s' - m'), row sum, bf16 P into the A-operand layout (masked / unloaded columns are zeros)
      float l = 0.f;
      for (int c = half; c < nch_all; c += 2) {
        uint32_t pk[16];
        if (c < nch_load) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
          float e[32];
          if (c < nch_full) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              e[j] = ex2_approx(fmaf(__uint_as_float(r[j]), scale2, -m));
              l += e[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = ex2_approx(fmaf(__uint_as_float(r[j]), scale2, -m));
              e[j] = (ok && c * 32 + j <= row) ? v : 0.f;
              l += e[j];
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        }
        // row rl of P block (c >> 1): 128-byte row, 16-byte chunks (c & 1) * 4 + j, XOR-swizzled with the row phase
        const uint32_t prow = smem_u32(((last - i) & 1) ? sP1 : sP) + (c >> 1) * AT_TILE_BYTES + (rl >> 3) * 1024 + (rl & 7) * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ch = ((c & 1) * 4 + j) ^ (rl & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (ch << 4)), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                       "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        }
      }
      s_l[half * 128 + rl] = l;
      fence_proxy_async_smem();     // generic-proxy writes of P -> visible to the tensor core
      tc_fence_before();            // our TMEM reads of S_i are done: O_i may overwrite its first 64 columns
      mbar_arrive(&bar_p[i]);
      AT_STAMP();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      lsum[i] = s_l[rl] + s_l[128 + rl];
    }
    if (last >= 1) epilogue(1, lsum[1]);
    epilogue(0, lsum[0]);
    AT_STAMP();
  }

  tc_fence_before();
  __syncthreads();
  AT_STAMP();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent form for the three-tile shapes (257 <= T <= 272: the VAS / VGGSound block size 265).  The one-head-per-CTA
// kernel above spends ~40 % of a CTA's life in fixed costs (launch and turnover, barrier set-up, TMEM allocation, waiting
// for its first loads, the last epilogues), and TMEM (496 of 512 columns) keeps a second CTA off the SM.  Here one CTA per
// SM walks over its heads with TWO Q | K | V buffers: the next head's operands land while the current head is computed,
// and the next head's S MMAs are issued as soon as the TMEM columns they overwrite have been consumed (S_2 of the next
// head after the epilogue of tile 2, S_1 / S_0 at the head boundary), so the softmax warps never wait for a load or for S.
// Shared memory (T = 265): 2 x (Q 34 KB | K 34 KB | V 34 KB) + one 16 KB block for the fifth P block of the long tile
// (P blocks 0..3 reuse the current buffer's Q | K, dead once its S MMAs are done) = 220 KB.
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_prefill_tc_persist_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm16,
                               const __grid_constant__ CUtensorMap tm32, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ptail = smem + 2 * p.buf_bytes;                      // P block 4 (keys 256..) of the long tile
  float* s_m = reinterpret_cast<float*>(ptail + AT_TILE_BYTES);  // [2][128]
  float* s_l = s_m + 256;                                        // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_l + 256);
  uint64_t* bar_qk = bars;         // [2] per buffer: Q_2 + K landed
  uint64_t* bar_q = bars + 2;      // [2] per buffer: Q_1 + Q_0 landed
  uint64_t* bar_v = bars + 4;      // [2] per buffer: V landed
  uint64_t* bar_s2 = bars + 6;     // per head: S_2 complete
  uint64_t* bar_sall = bars + 7;   // per head: S_1, S_0 complete
  uint64_t* bar_p = bars + 8;      // [3] per head: P_i written (256 arrivals)
  uint64_t* bar_o = bars + 11;     // [3] per head: O_i complete
  uint64_t* bar_oc = bars + 14;    // [3] per head: O_i read by every softmax thread (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x;
  const int n_my = (p.heads - static_cast<int>(blockIdx.x) + grid - 1) / grid;
  const int v_off = p.pk_bytes;                                  // V inside a head's buffer

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm);
    prefetch_tensormap(&tm16);
    prefetch_tensormap(&tm32);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_qk[i], 1);
      mbar_init(&bar_q[i], 1);
      mbar_init(&bar_v[i], 1);
    }
    mbar_init(bar_s2, 1);
    mbar_init(bar_sall, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bar_p[i], AT_THREADS - 64);
      mbar_init(&bar_o[i], 1);
      mbar_init(&bar_oc[i], AT_THREADS - 64);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int kfull = p.kmax / 128, krem = (p.kmax - kfull * 128) / 16;
      const uint32_t q0_bytes = p.q0_small ? 16 * 128 : AT_TILE_BYTES;
      auto issue_loads = [&](int it) {
        const int head = blockIdx.x + it * grid;
        const int b = head / p.nh, h = head - b * p.nh;
        uint8_t* buf = smem + (it & 1) * p.buf_bytes;
        const int bi = it & 1;
        mbar_arrive_expect_tx(&bar_qk[bi], AT_TILE_BYTES + static_cast<uint32_t>(p.kmax) * 128);
        for (int q = 0; q < 4; ++q)     // quarters of the long tile in reverse order (load balance of the softmax warps)
          tma_load_4d(buf + p.qoff[2] + q * 32 * 128, &tm32, &bar_qk[bi], h * GPT_HEAD_DIM, p.row0[2] + (3 - q) * 32, b, 0, kEvictFirst);
        for (int j = 0; j < kfull; ++j)
          tma_load_4d(buf + p.koff + j * AT_TILE_BYTES, &tm, &bar_qk[bi], p.C + h * GPT_HEAD_DIM, j * 128, b, 0, kEvictFirst);
        for (int j = 0; j < krem; ++j)
          tma_load_4d(buf + p.koff + (kfull * 128 + j * 16) * 128, &tm16, &bar_qk[bi], p.C + h * GPT_HEAD_DIM, kfull * 128 + j * 16, b, 0,
                      kEvictFirst);
        mbar_arrive_expect_tx(&bar_q[bi], AT_TILE_BYTES + q0_bytes);
        tma_load_4d(buf + p.qoff[1], &tm, &bar_q[bi], h * GPT_HEAD_DIM, p.row0[1], b, 0, kEvictFirst);
        tma_load_4d(buf + p.qoff[0], p.q0_small ? &tm16 : &tm, &bar_q[bi], h * GPT_HEAD_DIM, p.row0[0], b, 0, kEvictFirst);
        mbar_arrive_expect_tx(&bar_v[bi], static_cast<uint32_t>(p.kmax) * 128);
        for (int j = 0; j < kfull; ++j)
          tma_load_4d(buf + v_off + j * AT_TILE_BYTES, &tm, &bar_v[bi], 2 * p.C + h * GPT_HEAD_DIM, j * 128, b, 0, kEvictFirst);
        for (int j = 0; j < krem; ++j)
          tma_load_4d(buf + v_off + (kfull * 128 + j * 16) * 128, &tm16, &bar_v[bi], 2 * p.C + h * GPT_HEAD_DIM, kfull * 128 + j * 16, b,
                      0, kEvictFirst);
      };
      auto issue_s = [&](int it, int i) {
        uint8_t* buf = smem + (it & 1) * p.buf_bytes;
        const uint64_t da = make_smem_desc_sw128(smem_u32(buf + p.qoff[i]));
        for (int n0 = 0; n0 < p.kpad[i]; n0 += 256) {
          const int n = (p.kpad[i] - n0 < 256) ? p.kpad[i] - n0 : 256;
          const uint64_t db = make_smem_desc_sw128(smem_u32(buf + p.koff) + n0 * 128);
          const uint32_t idesc = at_idesc(128, n, 0);
#pragma unroll
          for (int k = 0; k < GPT_HEAD_DIM / 16; ++k)
            umma_bf16(tmem_base + p.scol[i] + n0, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
        }
      };
      const uint32_t idesc_pv = at_idesc(128, GPT_HEAD_DIM, 1);
      auto issue_pv = [&](int it, int i) {
        uint8_t* buf = smem + (it & 1) * p.buf_bytes;
        const int ksteps = p.kpad[i] / 16;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t pa = (ks < 16) ? smem_u32(buf) + (ks >> 2) * AT_TILE_BYTES : smem_u32(ptail);
          const uint64_t da = make_smem_desc_sw128(pa + (ks & 3) * 32);
          const uint64_t db = make_smem_desc_mn_sw128(smem_u32(buf + v_off) + ks * 2048, 8192, 1024);
          umma_bf16(tmem_base + p.ocol[i], da, db, idesc_pv, ks != 0 ? 1u : 0u);
        }
      };

      issue_loads(0);
      if (n_my > 1) issue_loads(1);
      mbar_wait(&bar_qk[0], 0);
      tc_fence_after();
      issue_s(0, 2);
      tc_commit(bar_s2);
      for (int it = 0; it < n_my; ++it) {
        const int bi = it & 1;
        const uint32_t ph = it & 1, lph = (it >> 1) & 1, pph = (it - 1) & 1;
        // S_1, S_0 of this head: their columns held S_1 / O_1 and S_0 of the previous head
        if (it > 0) mbar_wait(&bar_oc[1], pph);
        mbar_wait(&bar_q[bi], lph);
        tc_fence_after();
        issue_s(it, 1);
        issue_s(it, 0);
        tc_commit(bar_sall);
        // operands of the next head into the other buffer, once the previous head's last MMA has read it
        if (it >= 1 && it + 1 < n_my) {
          mbar_wait(&bar_o[0], pph);
          issue_loads(it + 1);
        }
        mbar_wait(&bar_v[bi], lph);
        mbar_wait(&bar_p[2], ph);
        tc_fence_after();
        issue_pv(it, 2);
        tc_commit(&bar_o[2]);
        mbar_wait(&bar_p[1], ph);
        tc_fence_after();
        issue_pv(it, 1);
        tc_commit(&bar_o[1]);
        if (it + 1 < n_my) {   // S_2 of the next head: O_2 of this one has been read, the next operands have landed
          mbar_wait(&bar_oc[2], ph);
          mbar_wait(&bar_qk[bi ^ 1], ((it + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(it + 1, 2);
          tc_commit(bar_s2);
        }
        mbar_wait(&bar_p[0], ph);
        tc_fence_after();
        issue_pv(it, 0);
        tc_commit(&bar_o[0]);
      }
    }
  } else if (warp >= 2) {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rl = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float scale2 = 1.4426950408889634f * 0.125f;
    auto tile_row = [&](int i) { return (i == 2) ? (3 - quarter) * 32 + lane : rl; };

    for (int it = 0; it < n_my; ++it) {
      const int head = blockIdx.x + it * grid;
      const int b = head / p.nh, h = head - b * p.nh;
      const uint32_t ph = it & 1;
      uint8_t* buf = smem + (it & 1) * p.buf_bytes;
      float lsum[3] = {0.f, 0.f, 0.f};

      auto epilogue = [&](int i, float l) {
        mbar_wait(&bar_o[i], ph);
        tc_fence_after();
        const int tr_ = tile_row(i);
        const bool ok = tr_ < p.nrows[i];
        const float inv = ok ? 1.0f / l : 0.f;
        __nv_bfloat16* dst = p.y + (static_cast<long long>(b) * p.T + p.row0[i] + tr_) * p.C + h * GPT_HEAD_DIM + half * 32;
        uint32_t r[32];
        tmem_ld_32x32(lane_base + p.ocol[i] + half * 32, r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bar_oc[i]);            // these TMEM columns may take the next head's S
        if (ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q;
            q.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
            q.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
            q.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
            q.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + j * 8) = q;
          }
        }
      };

      mbar_wait(bar_s2, ph);
      tc_fence_after();
      for (int i = 2; i >= 0; --i) {
        const int trow = tile_row(i);
        const int q0row = trow - lane;
        const bool ok = trow < p.nrows[i];
        const int row = p.row0[i] + trow;
        const int wlast_rl = (q0row + 31 < p.nrows[i]) ? q0row + 31 : p.nrows[i] - 1;
        const int wkeys = (wlast_rl >= q0row) ? p.row0[i] + wlast_rl + 1 : 0;
        const int nch_load = (wkeys + 31) / 32;
        const int nch_all = (p.kpad[i] + 31) / 32;
        const int nch_full = (q0row + 31 < p.nrows[i]) ? (p.row0[i] + q0row + 1) / 32 : 0;
        const uint32_t s_addr = lane_base + p.scol[i];
        if (i == 1) {                      // S_1, S_0 were issued at the head boundary
          mbar_wait(bar_sall, ph);
          tc_fence_after();
        }
        // pass 1: row maximum
        float m = -INFINITY;
        for (int c = half; c < nch_load; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
          if (c < nch_full) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) m = max3(m, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j <= row) m = fmaxf(m, __uint_as_float(r[j]));
          }
        }
        s_m[half * 128 + rl] = m;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        m = fmaxf(s_m[rl], s_m[128 + rl]) * scale2;
        // P_i reuses the region P_{i+1} was read from: that tile's P V must be complete -- its epilogue runs here.
        // The long tile writes over Q | K of this head: S_1 and S_0 must be complete.
        if (i == 2) mbar_wait(bar_sall, ph);
        else epilogue(i + 1, lsum[i + 1]);
        // pass 2
        float l = 0.f;
        for (int c = half; c < nch_all; c += 2) {
          uint32_t pk[16];
          if (c < nch_load) {
            uint32_t r[32];
            tmem_ld_32x32(s_addr + c * 32, r);
            tmem_ld_wait();
            float e[32];
            if (c < nch_full) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                e[j] = ex2_approx(fmaf(__uint_as_float(r[j]), scale2, -m));
                l += e[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = ex2_approx(fmaf(__uint_as_float(r[j]), scale2, -m));
                e[j] = (ok && c * 32 + j <= row) ? v : 0.f;
                l += e[j];
              }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          const uint32_t pblk = (c < 8) ? smem_u32(buf) + (c >> 1) * AT_TILE_BYTES : smem_u32(ptail);
          const uint32_t prow = pblk + (rl >> 3) * 1024 + (rl & 7) * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = ((c & 1) * 4 + j) ^ (rl & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (ch << 4)), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                         "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                         : "memory");
          }
        }
        s_l[half * 128 + rl] = l;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bar_p[i]);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        lsum[i] = s_l[rl] + s_l[128 + rl];
      }
      epilogue(0, lsum[0]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace

// true when attn_prefill_tc handles this shape (T rows in <= 3 query tiles whose score columns fit TMEM)
static bool attn_tc_plan(int T, AttnTcParams& p) {
  if (T < 1 || T > 3 * 128) return false;
  p.T = T;
  p.n_tiles = (T + 127) / 128;
  const int r = T - 128 * (p.n_tiles - 1);
  int col = 0;
  for (int i = 0; i < p.n_tiles; ++i) {
    p.row0[i] = (i == 0) ? 0 : r + 128 * (i - 1);
    p.nrows[i] = (i == 0) ? r : 128;
    p.kpad[i] = ((p.row0[i] + p.nrows[i] + 15) / 16) * 16;
    p.scol[i] = col;
    col += p.kpad[i];
  }
  p.ocol[0] = col;            // O_0 has its own 64 columns (S_0 may be narrower than that)
  col += 64;
  for (int i = 1; i < p.n_tiles; ++i) p.ocol[i] = p.scol[i];
  if (col > 512) return false;
  p.tmem_cols = 32;
  while (p.tmem_cols < col) p.tmem_cols *= 2;
  p.kmax = p.kpad[p.n_tiles - 1];
  p.q0_small = (p.nrows[0] <= 16) ? 1 : 0;
  // Q tiles last-first; a short tile 0 takes 2 KB (the MMA reads 128 rows from there: the tail is K data -- finite, and it
  // only reaches accumulator rows nobody reads unmasked)
  int off = 0;
  for (int i = p.n_tiles - 1; i >= 0; --i) {
    p.qoff[i] = off;
    off += (i == 0 && p.q0_small) ? 16 * 128 : AT_TILE_BYTES;
  }
  p.koff = off;
  const int qk = off + p.kmax * 128;
  const int pt = ((p.kmax + 63) / 64) * AT_TILE_BYTES;
  p.pk_bytes = ((qk > pt ? qk : pt) + 1023) & ~1023;
  if (p.pk_bytes < p.qoff[0] + AT_TILE_BYTES) p.pk_bytes = p.qoff[0] + AT_TILE_BYTES;   // tile 0's 128-row read stays inside
  p.p1_bytes = (p.n_tiles >= 2) ? ((p.kpad[p.n_tiles - 2] + 63) / 64) * AT_TILE_BYTES : 0;
  return true;
}

bool gpt_attention_prefill_tc_supported(int T) {
  static const bool off = getenv("MGV_ATTN_LEGACY") != nullptr;
  AttnTcParams p;
  return !off && attn_tc_plan(T, p);
}

int gpt_attention_prefill_tc(const __nv_bfloat16* qkv, int B, int T, int nh, __nv_bfloat16* y, cudaStream_t s, long long* trace) {
  AttnTcParams p;
  MGV_REQUIRE(attn_tc_plan(T, p), "attention (tcgen05): T=%d unsupported", T);
  if (B == 0) return MGV_OK;
  p.nh = nh;
  p.C = nh * GPT_HEAD_DIM;
  p.y = y;
  p.trace = trace;
  CUtensorMap tm, tm16, tm32;
  // qkv as (3C, T, B): a box never crosses into the next sequence, rows >= T are zero-filled
  MGV_TRY(make_tmap_nhwc_bf16(&tm, qkv, 3 * p.C, T, B, 1, GPT_HEAD_DIM, 128, 1, 1));
  MGV_TRY(make_tmap_nhwc_bf16(&tm16, qkv, 3 * p.C, T, B, 1, GPT_HEAD_DIM, 16, 1, 1));
  MGV_TRY(make_tmap_nhwc_bf16(&tm32, qkv, 3 * p.C, T, B, 1, GPT_HEAD_DIM, 32, 1, 1));
  const size_t smem = static_cast<size_t>(p.pk_bytes) + static_cast<size_t>(p.kmax) * 128 + p.p1_bytes + 2048 + 128 + 1024;
  static unsigned long long attr_mask = 0;   // per device
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
  }
  // three-tile shapes (T = 265): one persistent CTA per SM with double-buffered operands
  static const bool no_persist = getenv("MGV_ATTN_NO_PERSIST") != nullptr;
  AttnTcParams pp = p;
  pp.heads = B * nh;
  pp.pk_bytes = (p.koff + p.kmax * 128 + 1023) & ~1023;   // Q | K only: the fifth P block has its own 16 KB here
  pp.buf_bytes = pp.pk_bytes + p.kmax * 128;              // Q | K | V of one head
  const size_t smem_p = 2 * static_cast<size_t>(pp.buf_bytes) + AT_TILE_BYTES + 2048 + 256 + 1024;
  const bool persist = !no_persist && trace == nullptr && p.n_tiles == 3 && p.kmax > 256 && p.kmax <= 320 &&
                       4 * AT_TILE_BYTES <= pp.pk_bytes && p.qoff[0] + AT_TILE_BYTES <= pp.pk_bytes && smem_p <= 227 * 1024;
  if (persist) {
    static unsigned long long attr_mask_p = 0;   // per device
    if (first_use_on_this_device(attr_mask_p)) {
      MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_tc_persist_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
    }
    const int grid = pp.heads < num_sms() ? pp.heads : num_sms();
    attn_prefill_tc_persist_kernel<<<grid, AT_THREADS, smem_p, s>>>(tm, tm16, tm32, pp);
    MGV_CHECK_CUDA(cudaGetLastError());
    return MGV_OK;
  }
  attn_prefill_tc_kernel<<<B * nh, AT_THREADS, smem, s>>>(tm, tm16, tm32, p);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

}  // namespace mgv
