// Decode-step "stage program": one persistent kernel (one CTA per SM) walks a list of dependent stages
// (LayerNorm, swap-AB split-K tcgen05 GEMM, GELU) of one transformer block, separated by grid-wide barriers in
// global memory instead of kernel boundaries.  reference: transformer/minGPT.py:97-118 (Block.forward).
#pragma once
#include "mgv_sm100.cuh"

namespace mgv {

enum DecStageType { DST_LN = 0, DST_GEMM = 1 };
// GEMM epilogues (acc = sum_k act[b, k] * W[n, k]):
enum DecGemmMode {
  DGM_STORE_F32 = 0,   // out[b, n] = acc + bias[n]                       (fp32)
  DGM_ADD_F32 = 1,     // out[b, n] += acc + bias[n]    single writer     (fp32 residual stream)
  DGM_GELU_BF16 = 2,   // out[b, n] = bf16(gelu(acc + bias[n]))
  DGM_RED_F32 = 3      // out[b, n] += acc (+ bias[n] from split 0), split-K with reductions
};

constexpr int DP_MAX_KB = 16;   // 64-wide k blocks one CTA handles per GEMM stage (its K slice): K slice <= 1024

struct DecStage {
  int type;
  int B;                     // sequences (rows of every activation matrix), <= 64
  int C;                     // LN: row length
  // ---- DST_LN: ln_out[b, :] = bf16(LayerNorm(x[b, :]) * gamma + beta)
  const float* x;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* ln_out;
  // ---- DST_GEMM, swap-AB: a unit = 64 weight rows x 32 sequences x one K slice, everything resident in shared
  // memory (weights prefetched during the previous stages, activations loaded after the barrier)
  int map_w, map_x;          // indices into the tensor-map table (W: box 64 k x 64 rows, act: box 64 k x 32 rows)
  int ftiles;                // 64-row weight tiles
  int rhalves;               // 32-sequence groups (1 or 2)
  int splits;                // K slices; ftiles * rhalves * splits CTAs take part
  int nkb;                   // K / 64
  int n_feat;                // weight rows
  int mode;                  // DecGemmMode
  const float* bias;
  void* out;
  long long ldo;
};

// runs stages [s_begin, s_end) of `prog` (device memory) on n_ctas CTAs; `counter` is a zeroed device word used by
// this launch only.  trace (optional, 32 words): %globaltimer stamps of CTA 0 -- [0] entry, [1] dependency
// resolved, then per stage boundary [2+2i] CTA done / [3+2i] barrier passed, last = kernel end.
constexpr int DP_TRACE_WORDS = 32;
int decode_program_launch(const DecStage* prog, int s_begin, int s_end, const CUtensorMap* maps, unsigned int* counter,
                          int n_ctas, bool pdl, cudaStream_t s, unsigned long long* trace = nullptr);

}  // namespace mgv
