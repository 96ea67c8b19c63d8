// Probe (diagnostic, not product): tcgen05.mma kind::f16 (bf16) with MN-major operands in shared memory -- the form a
// wgrad GEMM dW = dY^T X needs to read dY [rows, out] and X [rows, in] without transposed copies.
// D[128 x 128] = A[128 x 64] * B[128 x 64]^T as four K = 16 MMAs.  Variants: which operands are MN-major, LBO / SBO roles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I melspec_gpt_vqvae_b200/csrc tools/probes/bf16_mn_probe.cu -o tools/probes/bf16_mn_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "mgv_sm100.cuh"
namespace mgv { void set_error(const char*, ...) {} const char* get_error() { return ""; } }
using namespace mgv;
using namespace mgv::sm100;

struct Variant { int a_mn, b_mn; unsigned lbo, sbo; };

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// element (row r of 128, k of 64) of an operand tile: K-major = rows of 128 B (64 k), 8-row groups 1024 B apart;
// MN-major = [block of 64 rows][k row of 128 B holding the 64 row-elements], blocks 8192 B apart
__device__ int off_kmajor(int r, int k) { return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2; }
__device__ int off_mnmajor(int r, int k) { return (r >> 6) * 8192 + k * 128 + (((((r & 63) >> 3) ^ (k & 7))) << 4) + (r & 7) * 2; }

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* D, Variant v) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + 16384;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  for (int i = t; i < 128 * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(sa + (v.a_mn ? off_mnmajor(r, k) : off_kmajor(r, k))) = A[r * 64 + k];
    *reinterpret_cast<__nv_bfloat16*>(sb + (v.b_mn ? off_mnmajor(r, k) : off_kmajor(r, k))) = Bm[r * 64 + k];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (t == 0) {
    const uint32_t idesc = idesc_bf16(128, 128, v.a_mn, v.b_mn);
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t da = v.a_mn ? make_smem_desc_mn_sw128(smem_u32(sa) + ks * 2048, v.lbo, v.sbo) : make_smem_desc_sw128(smem_u32(sa) + ks * 32);
      const uint64_t db = v.b_mn ? make_smem_desc_mn_sw128(smem_u32(sb) + ks * 2048, v.lbo, v.sbo) : make_smem_desc_sw128(smem_u32(sb) + ks * 32);
      umma_bf16(tb, da, db, idesc, ks != 0 ? 1u : 0u);
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
  std::vector<__nv_bfloat16> A(128 * 64), Bm(128 * 64);
  std::vector<float> Af(128 * 64), Bf(128 * 64), ref(128 * 128), D(128 * 128);
  srand(2);
  for (int i = 0; i < 128 * 64; ++i) { Af[i] = float(rand() % 17 - 8); Bf[i] = float(rand() % 13 - 6); A[i] = __float2bfloat16(Af[i]); Bm[i] = __float2bfloat16(Bf[i]); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      float s = 0;
      for (int k = 0; k < 64; ++k) s += Af[m * 64 + k] * Bf[n * 64 + k];
      ref[m * 128 + n] = s;
    }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, Bm.size() * 2); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bm.data(), Bm.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const Variant vs[] = {{0, 0, 0, 0}, {1, 0, 8192, 1024}, {1, 0, 1024, 8192}, {0, 1, 8192, 1024}, {1, 1, 8192, 1024}, {1, 1, 1024, 8192}};
  int vi = 0;
  for (const Variant& v : vs) {
    cudaMemset(dD, 0xff, D.size() * 4);
    probe<<<1, 128, 34 * 1024>>>(dA, dB, dD, v);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < 128 * 128; ++i) { double d = fabs(D[i] - ref[i]); if (!(d <= 1e-3)) ++bad; if (d > maxerr) maxerr = d; }
    printf("variant %d (a_mn=%d b_mn=%d lbo=%u sbo=%u): %s  bad=%d/16384 maxerr=%g  D[0][0..2]=%g %g %g ref=%g %g %g\n", vi++, v.a_mn, v.b_mn,
           v.lbo, v.sbo, cudaGetErrorString(e), bad, maxerr, D[0], D[1], D[2], ref[0], ref[1], ref[2]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
