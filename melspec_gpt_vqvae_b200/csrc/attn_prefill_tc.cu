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
// softmax: 4 warps, thread = query row (TMEM lane); exact two-pass form (row maximum, then exp2 / sum) over the causal
//          prefix of the row only (32-column chunks beyond a warp's last key are written as zeros, never loaded);
//          P goes to shared memory as bf16 in the K-major 128-byte-swizzled layout of an MMA A operand -- into the
//          Q | K region, which is dead once every S MMA has completed.
// O = P V: A = P (K-major), B = V in the layout TMA delivered it: rows = keys, 64 contiguous head dims = an MN-major
//          operand (descriptor form verified by tools/probes/bf16_mn_probe.cu: SBO = 8 keys x 128 B, K step = 2048 B).
// No [T, T] tensor reaches HBM.  The kernel covers the plain causal mask without the attention-map output and without
// the KV-cache fill; gpt_attention_prefill keeps the mma.sync kernel for those cases.
#include <stdlib.h>
#include "gpt_kernels.cuh"
#include "mgv_sm100.cuh"

namespace mgv {

using namespace sm100;

namespace {

constexpr int AT_THREADS = 192;     // warp 0: TMA + MMA issue, warp 1: TMEM allocation, warps 2..5: softmax / epilogue
constexpr int AT_TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int AT_MAX_TILES = 3;

struct AttnTcParams {
  int T, nh, C;
  int n_tiles;
  int row0[AT_MAX_TILES], nrows[AT_MAX_TILES], kpad[AT_MAX_TILES], scol[AT_MAX_TILES], ocol[AT_MAX_TILES];
  int nkb;            // 128-key boxes of K / V
  int tmem_cols;      // power of two
  __nv_bfloat16* y;   // [B*T, C]
};

// kind::f16, A = bf16 K-major, B = bf16 K-major (b_mn = 0) or MN-major (b_mn = 1), D = f32
__host__ __device__ constexpr uint32_t at_idesc(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tm, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [n_tiles] x 16 KB   } reused for P after the S MMAs
  uint8_t* sK = sQ + p.n_tiles * AT_TILE_BYTES;         // [nkb] x 16 KB       }
  const int pk_tiles = (p.n_tiles + p.nkb > (p.kpad[p.n_tiles - 1] + 63) / 64) ? p.n_tiles + p.nkb
                                                                               : (p.kpad[p.n_tiles - 1] + 63) / 64;
  uint8_t* sV = smem + pk_tiles * AT_TILE_BYTES;        // [nkb] x 16 KB: row = key, 128 B = 64 head dims
  uint8_t* sP = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + p.nkb * AT_TILE_BYTES);
  uint64_t* bar_qk = bars;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;                           // [3] P_i written
  uint64_t* bar_o = bars + 6;                           // [3] O_i complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.nh, h = blockIdx.x - b * p.nh;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    for (int i = 0; i < AT_MAX_TILES; ++i) {
      mbar_init(&bar_p[i], 128);
      mbar_init(&bar_o[i], 1);
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
      // ---- loads: Q tiles + K boxes on one barrier, V boxes on another (V is needed only after the first softmax)
      mbar_arrive_expect_tx(bar_qk, static_cast<uint32_t>(p.n_tiles + p.nkb) * AT_TILE_BYTES);
      for (int i = 0; i < p.n_tiles; ++i)
        tma_load_4d(sQ + i * AT_TILE_BYTES, &tm, bar_qk, h * GPT_HEAD_DIM, p.row0[i], b, 0, kEvictFirst);
      for (int j = 0; j < p.nkb; ++j)
        tma_load_4d(sK + j * AT_TILE_BYTES, &tm, bar_qk, p.C + h * GPT_HEAD_DIM, j * 128, b, 0, kEvictFirst);
      mbar_arrive_expect_tx(bar_v, static_cast<uint32_t>(p.nkb) * AT_TILE_BYTES);
      for (int j = 0; j < p.nkb; ++j)
        tma_load_4d(sV + j * AT_TILE_BYTES, &tm, bar_v, 2 * p.C + h * GPT_HEAD_DIM, j * 128, b, 0, kEvictFirst);

      // ---- S_i = Q_i K^T (all tiles back to back; N in chunks of <= 256 keys, 4 K steps of 16 head dims)
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      for (int i = 0; i < p.n_tiles; ++i) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(sQ + i * AT_TILE_BYTES));
        for (int n0 = 0; n0 < p.kpad[i]; n0 += 256) {
          const int n = (p.kpad[i] - n0 < 256) ? p.kpad[i] - n0 : 256;
          const uint64_t db = make_smem_desc_sw128(smem_u32(sK) + n0 * 128);
          const uint32_t idesc = at_idesc(128, n, 0);
#pragma unroll
          for (int k = 0; k < GPT_HEAD_DIM / 16; ++k)
            umma_bf16(tmem_base + p.scol[i] + n0, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
        }
      }
      tc_commit(bar_s);

      // ---- O_i = P_i V as soon as the softmax warps have written P_i
      mbar_wait(bar_v, 0);
      const uint32_t idesc_pv = at_idesc(128, GPT_HEAD_DIM, 1);
      for (int i = 0; i < p.n_tiles; ++i) {
        mbar_wait(&bar_p[i], 0);
        tc_fence_after();
        const int ksteps = p.kpad[i] / 16;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = make_smem_desc_sw128(smem_u32(sP) + (ks >> 2) * AT_TILE_BYTES + (ks & 3) * 32);
          const uint64_t db = make_smem_desc_mn_sw128(smem_u32(sV) + ks * 2048, 8192, 1024);
          umma_bf16(tmem_base + p.ocol[i], da, db, idesc_pv, ks != 0 ? 1u : 0u);
        }
        tc_commit(&bar_o[i]);
      }
    }
  } else if (warp >= 2) {
    // ---- softmax + epilogue: TMEM lane quarter = warp & 3, thread = one query row of the tile
    const int quarter = warp & 3;
    const int rl = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // 1/sqrt(d) (minGPT.py:81) folded with log2(e): probabilities are exp2(s' - max')
    const float scale2 = 1.4426950408889634f * 0.125f;
    mbar_wait(bar_s, 0);
    tc_fence_after();

    float l_prev = 0.f;
    auto epilogue = [&](int i, float l) {
      mbar_wait(&bar_o[i], 0);
      tc_fence_after();
      const bool ok = rl < p.nrows[i];
      const float inv = ok ? 1.0f / l : 0.f;
      __nv_bfloat16* dst = p.y + (static_cast<long long>(b) * p.T + p.row0[i] + rl) * p.C + h * GPT_HEAD_DIM;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(lane_base + p.ocol[i] + c * 32, r);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q;
            q.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
            q.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
            q.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
            q.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c * 32 + j * 8) = q;
          }
        }
      }
    };

    for (int i = 0; i < p.n_tiles; ++i) {
      const bool ok = rl < p.nrows[i];
      const int row = p.row0[i] + rl;                       // keys 0..row are visible (tril mask, minGPT.py:65-68, :82)
      // last key any row of this warp can see -> number of 32-column chunks this warp has to load (warp-uniform)
      const int wlast_rl = (quarter * 32 + 31 < p.nrows[i]) ? quarter * 32 + 31 : p.nrows[i] - 1;
      const int wkeys = (wlast_rl >= quarter * 32) ? p.row0[i] + wlast_rl + 1 : 0;
      const int nch_load = (wkeys + 31) / 32;
      const int nch_all = (p.kpad[i] + 31) / 32;
      const uint32_t s_addr = lane_base + p.scol[i];

      // pass 1: exact row maximum of the scaled scores
      float m = -INFINITY;
      for (int c = 0; c < nch_load; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(s_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j <= row) m = fmaxf(m, __uint_as_float(r[j]));
      }
      m *= scale2;
      // the previous tile's output: its P V MMAs ran while pass 1 was reading S_i; they must be complete before P_i
      // overwrites the shared-memory region they read from
      if (i > 0) epilogue(i - 1, l_prev);

      // pass 2: p = exp2(s' - m'), row sum, bf16 P into the A-operand layout (masked / unloaded columns are zeros)
      float l = 0.f;
      for (int c = 0; c < nch_all; ++c) {
        uint32_t pk[16];
        if (c < nch_load) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
          float e[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = ex2_approx(fmaf(__uint_as_float(r[j]), scale2, -m));
            e[j] = (ok && c * 32 + j <= row) ? v : 0.f;
            l += e[j];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        }
        // row rl of P block (c >> 1): 128-byte row, 16-byte chunks (c & 1) * 4 + j, XOR-swizzled with the row phase
        const uint32_t prow = smem_u32(sP) + (c >> 1) * AT_TILE_BYTES + (rl >> 3) * 1024 + (rl & 7) * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ch = ((c & 1) * 4 + j) ^ (rl & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (ch << 4)), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                       "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();     // generic-proxy writes of P -> visible to the tensor core
      tc_fence_before();            // our TMEM reads of S_i are done: O_i may overwrite its first 64 columns
      mbar_arrive(&bar_p[i]);
      l_prev = l;
    }
    epilogue(p.n_tiles - 1, l_prev);
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
  p.nkb = (p.kpad[p.n_tiles - 1] + 127) / 128;
  return true;
}

bool gpt_attention_prefill_tc_supported(int T) {
  static const bool off = getenv("MGV_ATTN_LEGACY") != nullptr;
  AttnTcParams p;
  return !off && attn_tc_plan(T, p);
}

int gpt_attention_prefill_tc(const __nv_bfloat16* qkv, int B, int T, int nh, __nv_bfloat16* y, cudaStream_t s) {
  AttnTcParams p;
  MGV_REQUIRE(attn_tc_plan(T, p), "attention (tcgen05): T=%d unsupported", T);
  if (B == 0) return MGV_OK;
  p.nh = nh;
  p.C = nh * GPT_HEAD_DIM;
  p.y = y;
  CUtensorMap tm;
  // qkv as (3C, T, B): a box never crosses into the next sequence, rows >= T are zero-filled
  MGV_TRY(make_tmap_nhwc_bf16(&tm, qkv, 3 * p.C, T, B, 1, GPT_HEAD_DIM, 128, 1, 1));
  const int p_tiles = (p.kpad[p.n_tiles - 1] + 63) / 64;
  const int pk_tiles = p.n_tiles + p.nkb > p_tiles ? p.n_tiles + p.nkb : p_tiles;
  const size_t smem = static_cast<size_t>(pk_tiles + p.nkb) * AT_TILE_BYTES + 128 + 1024;
  static unsigned long long attr_mask = 0;   // per device
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MGV_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
  }
  attn_prefill_tc_kernel<<<B * nh, AT_THREADS, smem, s>>>(tm, p);
  MGV_CHECK_CUDA(cudaGetLastError());
  return MGV_OK;
}

}  // namespace mgv
