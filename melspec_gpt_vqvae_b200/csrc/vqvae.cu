#include "vqvae.cuh"
namespace mgv {
struct Vqvae { int x; };
int vqvae_create(int, int, Vqvae**) { set_error("stub"); return MGV_ERR_UNSUPPORTED; }
int vqvae_destroy(Vqvae*) { return 0; }
int vqvae_load_weight(Vqvae*, const char*, const float*, long long, cudaStream_t) { return MGV_ERR_UNSUPPORTED; }
int vqvae_decode(Vqvae*, const long long*, const float*, int, float*, cudaStream_t) { return MGV_ERR_UNSUPPORTED; }
int vqvae_encode(Vqvae*, const float*, int, float*, cudaStream_t) { return MGV_ERR_UNSUPPORTED; }
long long vqvae_last_launches(const Vqvae*) { return 0; }
}
