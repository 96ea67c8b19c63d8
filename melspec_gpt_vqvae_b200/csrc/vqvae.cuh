// VQVAE encoder / decoder handle (see vqvae.cu).
#pragma once
#include "mgv_common.cuh"

namespace mgv {

struct Vqvae;

int vqvae_create(int num_embeddings, int embedding_dim, Vqvae** out);
int vqvae_destroy(Vqvae* v);
int vqvae_load_weight(Vqvae* v, const char* name, const float* src, long long numel, cudaStream_t s);
// exactly one of idx (B, 265 codes, row-major grid) / quant_bchw (B,256,5,53) is non-null
int vqvae_decode(Vqvae* v, const long long* idx, const float* quant_bchw, int B, float* mel_out, cudaStream_t s);
int vqvae_encode(Vqvae* v, const float* mel, int B, float* z_out, cudaStream_t s);
long long vqvae_last_launches(const Vqvae* v);

}  // namespace mgv
