// Fused codebook kernels (see vq.cu).
#pragma once
#include "mgv_common.cuh"

namespace mgv {

int vq_argmin(const float* z, const float* codebook, int B, int D, int HW, int K, long long* idx_out, float* dmin_out,
              cudaStream_t stream);
int vq_finish(const float* z, const float* codebook, const long long* idx, int B, int D, int HW, int K,
              float commitment_cost, float* quantized, float* encodings, float* loss_out, float* perplexity_out,
              void* workspace, cudaStream_t stream);
int vq_gather(const long long* idx, const float* codebook, long long n_vec, int D, int HW, int K, float* out,
              int* bad_index_flag, cudaStream_t stream);

}  // namespace mgv
