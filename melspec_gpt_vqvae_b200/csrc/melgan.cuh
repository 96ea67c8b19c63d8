// MelGAN generator handle (melgan.cu).  reference: vocoder/modules.py:38-80
#pragma once
#include <cuda_runtime.h>

namespace mgv {

struct Melgan;
int melgan_create(Melgan** out, int n_mel, int ngf, int n_res);
int melgan_destroy(Melgan* m);
int melgan_load_weight(Melgan* m, const char* name, const float* src, long long numel, cudaStream_t s);
int melgan_reset_biases(Melgan* m, cudaStream_t s);
int melgan_forward(Melgan* m, const float* mel, int B, int T, float* wave, cudaStream_t s);
int melgan_last_launches(const Melgan* m);

}  // namespace mgv
