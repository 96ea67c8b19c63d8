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
int num_sms();


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
// x * sigmoid(x); fast division (rcp.approx + mul): the result is rounded to bf16 by every caller that stores it
__device__ __forceinline__ float swish(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

#endif  // __CUDACC__

}  // namespace mgv
