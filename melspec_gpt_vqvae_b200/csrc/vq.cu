// Fused codebook kernels.
//   reference: VectorQuantizer.forward            vqvae/big_model_attn_gan.py:19-54
//              VectorQuantizer.get_codebook_entry vqvae/big_model_attn_gan.py:56-71
//
// vq_argmin_kernel: for every latent vector x (read straight from the BCHW tensor, no
// permute copy) computes, in fp32 and in the reference's operation order,
//     d_j = fl( fl(|x|^2 + |e_j|^2) - fl(2 * <x, e_j>) )          (reference :28-30)
// and the first index attaining the minimum (torch.argmin semantics, :33).
// Every sum is a sequential fmaf chain over the channel index (k = 0..D-1, starting from
// +0.0f); oracle/vq_oracle.c restates exactly this, so indices AND distances are
// bit-reproducible on the CPU.
//
// Layout: one CTA owns 128 consecutive vectors x all K<=128 codes; the codebook is staged
// once per (persistent) CTA in shared memory, transposed to [k][code]; x is streamed in
// 64-channel chunks with cp.async (double buffered) as [k][vector].  Each of the 256
// threads keeps an 8 vector x 8 code fp32 accumulator tile; the argmin over the 16 lanes
// that share a vector is a warp-shuffle reduction with lowest-index tie-break.
#include "mgv_common.cuh"

namespace mgv {

namespace {

constexpr int VQ_TILE_V = 128;   // vectors per CTA tile
constexpr int VQ_MAX_K = 128;    // codes held in the register tiling
constexpr int VQ_KC = 64;        // channels per streamed chunk
constexpr int VQ_THREADS = 256;

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 4 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// dynamic smem: es[D][128] | ee[128] | xs[2][VQ_KC][128]
__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_argmin_kernel(const float* __restrict__ z, const float* __restrict__ codebook, int B, int D, int HW, int K,
                 long long* __restrict__ idx_out, float* __restrict__ dmin_out, int code0, int merge) {
  // Codebooks larger than the 128-code register tiling run as passes over chunks of 128 codes (code0 = first code of
  // this pass, ascending): a pass with merge = 1 keeps the earlier passes' winner unless its own minimum is strictly
  // smaller, which preserves "lowest index wins" across chunks.  K counts the codes of THIS chunk.
  codebook += static_cast<size_t>(code0) * D;
  extern __shared__ __align__(16) float vq_smem[];
  float* es = vq_smem;                        // [D][128]
  float* ee = es + static_cast<size_t>(D) * VQ_MAX_K;  // [128]
  float* xs = ee + VQ_MAX_K;                  // [2][VQ_KC][128]

  const int t = threadIdx.x;
  const long long N = static_cast<long long>(B) * HW;
  const int num_tiles = static_cast<int>((N + VQ_TILE_V - 1) / VQ_TILE_V);

  // ---- stage the codebook transposed: es[k][code]; unused codes get +inf distance later
  {
    const int code = t & 127;
    for (int k4 = (t >> 7); k4 < D / 4; k4 += 2) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (code < K) v = __ldg(reinterpret_cast<const float4*>(codebook + static_cast<size_t>(code) * D) + k4);
      es[(4 * k4 + 0) * VQ_MAX_K + code] = v.x;
      es[(4 * k4 + 1) * VQ_MAX_K + code] = v.y;
      es[(4 * k4 + 2) * VQ_MAX_K + code] = v.z;
      es[(4 * k4 + 3) * VQ_MAX_K + code] = v.w;
    }
  }
  __syncthreads();
  if (t < VQ_MAX_K) {
    float s = 0.f;
    for (int k = 0; k < D; ++k) {
      const float e = es[k * VQ_MAX_K + t];
      s = fmaf(e, e, s);
    }
    ee[t] = s;
  }
  // (visibility of ee is covered by the first __syncthreads of the tile loop)

  const int cg = t & 15;        // code group: codes cg*8 .. cg*8+7
  const int vg = t >> 4;        // vector group: vectors vg*8 .. vg*8+7
  const int lv = t & 127;       // vector this thread streams in
  const int lk = t >> 7;        // 0/1: even / odd channels of the chunk
  const int nchunks = D / VQ_KC;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long n0 = static_cast<long long>(tile) * VQ_TILE_V;
    const long long nl = n0 + lv;
    const bool lvalid = nl < N;
    const long long lb = lvalid ? nl / HW : 0;
    const int lpos = lvalid ? static_cast<int>(nl - lb * HW) : 0;
    const float* gsrc = z + (lb * D) * HW + lpos;  // + k*HW

    auto issue_chunk = [&](int kc, int buf) {
      float* dst = xs + buf * (VQ_KC * VQ_TILE_V);
#pragma unroll 8
      for (int i = 0; i < VQ_KC / 2; ++i) {
        const int kl = lk + 2 * i;
        cp_async_f32(dst + kl * VQ_TILE_V + lv, gsrc + static_cast<long long>(kc * VQ_KC + kl) * HW, lvalid);
      }
      cp_async_commit();
    };

    // packed fp32x2 accumulators (fma.rn.f32x2 = two independent IEEE fmas, bit-identical to scalar fmaf):
    // acc2[ip][j] = dot products of vectors (2*ip, 2*ip+1) with code j
    float2 acc2[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc2[i][j] = make_float2(0.f, 0.f);
    float xx_self = 0.f;  // |x|^2 of vector vg*8 + (cg & 7)

    issue_chunk(0, 0);
    for (int kc = 0; kc < nchunks; ++kc) {
      const int buf = kc & 1;
      if (kc + 1 < nchunks) {
        issue_chunk(kc + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const float* xb = xs + buf * (VQ_KC * VQ_TILE_V);
      const float* eb = es + static_cast<size_t>(kc) * VQ_KC * VQ_MAX_K;
#pragma unroll 4
      for (int k = 0; k < VQ_KC; ++k) {
        const float4 x0 = *reinterpret_cast<const float4*>(xb + k * VQ_TILE_V + vg * 8);
        const float4 x1 = *reinterpret_cast<const float4*>(xb + k * VQ_TILE_V + vg * 8 + 4);
        const float4 e0 = *reinterpret_cast<const float4*>(eb + k * VQ_MAX_K + cg * 8);
        const float4 e1 = *reinterpret_cast<const float4*>(eb + k * VQ_MAX_K + cg * 8 + 4);
        const float xself = xb[k * VQ_TILE_V + vg * 8 + (cg & 7)];
        const float2 xp[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y),
                              make_float2(x1.z, x1.w)};
        const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 e2 = make_float2(ev[j], ev[j]);
#pragma unroll
          for (int ip = 0; ip < 4; ++ip) acc2[ip][j] = __ffma2_rn(xp[ip], e2, acc2[ip][j]);
        }
        xx_self = fmaf(xself, xself, xx_self);
      }
      __syncthreads();  // everyone done with xs[buf] before it is refilled
    }

    // ---- distances + argmin
    float eev[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) eev[j] = ee[cg * 8 + j];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // |x|^2 of vector vg*8+i lives in the lane of this half-warp with (cg & 7) == i
      const float xx = __shfl_sync(0xffffffffu, xx_self, (threadIdx.x & 16) | i);
      float best = __int_as_float(0x7f800000);  // +inf
      int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int code = cg * 8 + j;
        const float dot = (i & 1) ? acc2[i >> 1][j].y : acc2[i >> 1][j].x;
        const float d = __fsub_rn(__fadd_rn(xx, eev[j]), __fmul_rn(2.0f, dot));
        if (code < K && (d < best || bi == 0x7fffffff)) {  // strict <: first index wins inside the thread
          best = d;
          bi = code;
        }
      }
      // reduce over the 16 lanes (code groups) of this half-warp: min value, then min index
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      const long long n = n0 + vg * 8 + i;
      if (cg == 0 && n < N) {
        const bool keep_old = merge && !(best < dmin_out[n]);
        if (!keep_old) {
          idx_out[n] = code0 + bi;
          if (dmin_out) dmin_out[n] = best;
        }
      }
    }
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

int vq_argmin(const float* z, const float* codebook, int B, int D, int HW, int K, long long* idx_out, float* dmin_out,
              cudaStream_t stream) {
  MGV_REQUIRE(B >= 0 && HW > 0, "vq_argmin: B=%d HW=%d", B, HW);
  if (B == 0) return MGV_OK;   // empty batch: nothing to do (pointers may be null)
  MGV_REQUIRE(z && codebook && idx_out, "vq_argmin: null pointer");
  MGV_REQUIRE(K >= 1 && K <= 65536, "vq_argmin: num_embeddings=%d unsupported (1..65536)", K);
  MGV_REQUIRE(D >= VQ_KC && D % VQ_KC == 0 && D <= 256, "vq_argmin: embedding_dim=%d must be 64, 128, 192 or 256", D);
  const long long N = static_cast<long long>(B) * HW;
  const int num_tiles = static_cast<int>((N + VQ_TILE_V - 1) / VQ_TILE_V);
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  const size_t smem = (static_cast<size_t>(D) * VQ_MAX_K + VQ_MAX_K + 2 * VQ_KC * VQ_TILE_V) * sizeof(float);
  static unsigned long long attr_mask = 0;   // per device (and per template instantiation)
  if (first_use_on_this_device(attr_mask)) {
    MGV_CHECK_CUDA(cudaFuncSetAttribute(vq_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  // codebooks beyond one 128-code pass: the passes hand the running minimum to each other through dmin_out
  // (caller-owned, so that concurrent streams / devices never share scratch memory)
  MGV_REQUIRE(K <= VQ_MAX_K || dmin_out != nullptr, "vq_argmin: num_embeddings=%d > %d needs a dmin buffer of B*HW floats", K, VQ_MAX_K);
  for (int code0 = 0; code0 < K; code0 += VQ_MAX_K) {
    const int kc = (K - code0 < VQ_MAX_K) ? K - code0 : VQ_MAX_K;
    vq_argmin_kernel<<<grid, VQ_THREADS, smem, stream>>>(z, codebook, B, D, HW, kc, idx_out, dmin_out, code0, code0 > 0 ? 1 : 0);
    MGV_CHECK_CUDA(cudaGetLastError());
  }
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
