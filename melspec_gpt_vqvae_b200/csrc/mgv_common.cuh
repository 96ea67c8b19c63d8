// Shared host/device helpers for libmgv (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/mgv.h"  // MGV_OK / MGV_ERR_* codes

#if defined(__CUDA_ARCH__) && !(defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM103_ALL))
#error "libmgv is written for sm_100a only: compile with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace mgv {

// ---------------------------------------------------------------- errors
// 0 = ok; nonzero = error (MGV_ERR_* from include/mgv.h); message via mgv_last_error() (thread-local).
void set_error(const char* fmt, ...);
const char* get_error();

#define MGV_CHECK_CUDA(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      mgv::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                     cudaGetErrorString(_e));                                        \
      return MGV_ERR_CUDA;                                                      \
    }                                                                                \
  } while (0)

#define MGV_REQUIRE(cond, ...)                                                      \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      mgv::set_error(__VA_ARGS__);                                                   \
      return MGV_ERR_INVALID;                                                   \
    }                                                                                \
  } while (0)

#define MGV_TRY(expr)                                                               \
  do {                                                                               \
    int _r = (expr);                                                                 \
    if (_r != 0) return _r;                                                          \
  } while (0)

int check_device();      // MGV_OK iff the current device is compute capability 10.x
int num_sms();           // SM count of the CURRENT device (cached per device)
// true the first time it is called for `mask` on the current device: per-device one-off setup such as
// cudaFuncSetAttribute (function attributes are per device, and a process may drive several GPUs)
bool first_use_on_this_device(unsigned long long& mask);


static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// launch helper: optional programmatic dependent launch (PDL) attribute
struct LaunchCfg {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl) {
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cfg.attrs = attrs;
    cfg.numAttrs = 0;
    if (pdl) {
      attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attrs[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.numAttrs = 1;
    }
  }
};

// Statistics + fold vectors of a FOLD_LN GEMM, handed to whoever reads its raw accumulator:
//   value[f] = rstd * (acc[f] - mu * sw[f]) + bp[f],   mu / rstd from sum_z stats[z * stride + row] over `dim` elements.
struct LnFold {
  const float2* stats = nullptr;   // [nparts][stride] per-K-slice partial (sum, sum of squares) of the LayerNorm input rows
  int nparts = 0;
  int stride = 0;
  int dim = 0;                     // row length of the LayerNorm (n_embd)
  const float* sw = nullptr;       // [features] W gamma
  const float* bp = nullptr;       // [features] W beta + bias
};

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}

// programmatic dependent launch: wait for the upstream grid's memory to be visible /
// allow the downstream grid to start its prologue.  Both are no-ops when the kernel was
// launched without the PDL attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// exact (erf) GELU as torch.nn.GELU() default (reference: transformer/minGPT.py:102)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// The same GELU with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 plus the approx-unit error of rcp / ex2,
// ~1e-6, before the caller rounds to bf16 whose half-ulp is 2e-3 relative): 16 branch-free instructions, two of them
// on the special-function unit, instead of erff's two divergent polynomial paths.  Used where the activation is
// recomputed by several CTAs (gemm_decode_fold.cu).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float erf_abs = fmaf(-p * t, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);      // 0.5 x (1 + sign(x) erf|z|)
}
// x * sigmoid(x); fast division (rcp.approx + mul): the result is rounded to bf16 by every caller that stores it
__device__ __forceinline__ float swish(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

#endif  // __CUDACC__

}  // namespace mgv
