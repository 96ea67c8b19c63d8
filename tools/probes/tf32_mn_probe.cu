// Probe (diagnostic, not product): semantics of tcgen05.mma kind::tf32 with an MN-major A operand in shared memory.
// One CTA, D[128 x 128] = A[128 x 32] * B[128 x 32]^T as four K = 8 MMAs; A is stored by plain stores in candidate
// layouts, B K-major with the 128-byte swizzle.  Prints the maximum error of each variant against the CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I melspec_gpt_vqvae_b200/csrc tools/probes/tf32_mn_probe.cu -o /tmp/probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "mgv_sm100.cuh"
namespace mgv { void set_error(const char*, ...) {} const char* get_error() { return ""; } }
using namespace mgv;
using namespace mgv::sm100;

struct Variant { int a_mn; int swz; unsigned lbo, sbo; int kstep_bytes; int layout; };

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* Bm, float* D, Variant v) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;            // 16 KB
  uint8_t* sb = smem + 16384;    // 16 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  // B: K-major SW128: row n (128 B = 32 floats), 8-row groups 1024 B apart
  for (int i = t; i < 128 * 32; i += 128) {
    const int n = i >> 5, k = i & 31;
    const int off = (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 2) ^ (n & 7))) << 4) + (k & 3) * 4;
    *reinterpret_cast<float*>(sb + off) = Bm[n * 32 + k];
  }
  for (int i = t; i < 128 * 32; i += 128) {
    const int m = i & 127, k = i >> 7;
    int off;
    if (v.layout == 0) {          // MN-major: [kg][mb][r][128 B], chunk swizzled by r
      const int c16 = (m & 31) >> 2;
      off = (k >> 3) * 4096 + (m >> 5) * 1024 + (k & 7) * 128 + (((v.swz ? (c16 ^ (k & 7)) : c16)) << 4) + (m & 3) * 4;
    } else if (v.layout == 1) {   // MN-major: [mb][kg][r][128 B]
      const int c16 = (m & 31) >> 2;
      off = (m >> 5) * 4096 + (k >> 3) * 1024 + (k & 7) * 128 + (((v.swz ? (c16 ^ (k & 7)) : c16)) << 4) + (m & 3) * 4;
    } else {                      // K-major (like B)
      off = (m >> 3) * 1024 + (m & 7) * 128 + ((((k >> 2) ^ (m & 7))) << 4) + (k & 3) * 4;
    }
    *reinterpret_cast<float*>(sa + off) = A[m * 32 + k];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (t == 0) {
    const uint32_t idesc = make_idesc_tf32_f32(128, 128, v.a_mn, 0);
    for (int kg = 0; kg < 4; ++kg) {
      uint64_t da;
      if (v.layout == 2) da = make_smem_desc_sw128(smem_u32(sa) + kg * 32);
      else da = make_smem_desc_mn_sw128(smem_u32(sa) + kg * v.kstep_bytes, v.lbo, v.sbo);
      if (!v.swz && v.layout != 2) da &= ~(7ull << 61);   // SWIZZLE_NONE
      const uint64_t db = make_smem_desc_sw128(smem_u32(sb) + kg * 32);
      umma_tf32(tb, da, db, idesc, kg != 0 ? 1u : 0u);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(tb + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[t * 128 + ch * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 128); }
}

int main() {
  std::vector<float> A(128 * 32), Bm(128 * 32), ref(128 * 128), D(128 * 128);
  srand(1);
  for (auto& x : A) x = float(rand() % 17 - 8);
  for (auto& x : Bm) x = float(rand() % 13 - 6);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      float s = 0;
      for (int k = 0; k < 32; ++k) s += A[m * 32 + k] * Bm[n * 32 + k];
      ref[m * 128 + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bm.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bm.data(), Bm.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const Variant vs[] = {
      {1, 1, 1024, 4096, 4096, 0},   // as written in vq.cu
      {1, 1, 4096, 1024, 4096, 0},   // LBO / SBO swapped
      {1, 1, 4096, 1024, 1024, 1},   // [mb][kg] order: LBO = 4096 between M blocks
      {1, 1, 1024, 4096, 1024, 1},
      {0, 1, 0, 0, 0, 2},            // K-major A (sanity of kind::tf32 itself)
      {1, 0, 1024, 4096, 4096, 0},   // no swizzle
      {1, 0, 4096, 1024, 4096, 0},
      {1, 0, 128, 1024, 4096, 0},
  };
  int vi = 0;
  for (const Variant& v : vs) {
    cudaMemset(dD, 0xff, D.size() * 4);
    probe<<<1, 128, 34 * 1024>>>(dA, dB, dD, v);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < 128 * 128; ++i) { double d = fabs(D[i] - ref[i]); if (!(d <= 1e-3)) ++bad; if (d > maxerr) maxerr = d; }
    printf("variant %d (a_mn=%d swz=%d lbo=%u sbo=%u kstep=%d layout=%d): %s  bad=%d/16384 maxerr=%g   D[0][0..3]=%g %g %g %g  ref=%g %g %g %g\n",
           vi++, v.a_mn, v.swz, v.lbo, v.sbo, v.kstep_bytes, v.layout, cudaGetErrorString(e), bad, maxerr, D[0], D[1], D[2], D[3],
           ref[0], ref[1], ref[2], ref[3]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
