// C ABI (include/mgv.h) + process-wide helpers (errors, device check, tensor maps).
#include <cudaTypedefs.h>
#include <exception>
#include "gemm_tc.cuh"
#include "vqvae_kernels.cuh"
#include "gpt_kernels.cuh"
#include "gpt.cuh"
#include "gpt_train.cuh"
#include "melgan.cuh"
#include "vq.cuh"
#include "vqvae.cuh"

namespace mgv {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int check_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (libmgv has no CPU fallback)", cudaGetErrorString(e));
    cudaGetLastError();
    return MGV_ERR_DEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("device %d is sm_%d%d; libmgv is built for sm_100a (B200) only", dev, major, minor);
    return MGV_ERR_DEVICE;
  }
  return MGV_OK;
}

int num_sms() {
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& slot = n[dev & 63];
  if (slot == 0) {
    cudaDeviceGetAttribute(&slot, cudaDevAttrMultiProcessorCount, dev);
    if (slot <= 0) slot = 148;
  }
  return slot;
}

bool first_use_on_this_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// ------------------------------------------------------------------ tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  auto enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MGV_ERR_CUDA;
  }
  MGV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p not 16-byte aligned", base);
  MGV_REQUIRE(row_stride_bytes % 16 == 0, "tensor map: row stride %llu not a multiple of 16 bytes",
              static_cast<unsigned long long>(row_stride_bytes));
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d %llux%llu box %ux%u) failed: CUresult %d",
              static_cast<unsigned long long>(inner), static_cast<unsigned long long>(outer), box_inner, box_outer,
              static_cast<int>(r));
    return MGV_ERR_CUDA;
  }
  return MGV_OK;
}

// fp32 matrix [outer rows][inner] for kind::tf32 operands: 128-byte swizzle = 32 floats per row segment
int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
  auto enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MGV_ERR_CUDA;
  }
  MGV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p not 16-byte aligned", base);
  MGV_REQUIRE(row_stride_bytes % 16 == 0 && box_inner * 4 == 128, "tensor map (f32): row stride %llu / box %u",
              static_cast<unsigned long long>(row_stride_bytes), box_inner);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d f32 %llux%llu box %ux%u) failed: CUresult %d",
              static_cast<unsigned long long>(inner), static_cast<unsigned long long>(outer), box_inner, box_outer,
              static_cast<int>(r));
    return MGV_ERR_CUDA;
  }
  return MGV_OK;
}

int make_tmap_nhwc_bf16(CUtensorMap* out, const void* base, int C, int W, int H, int N, uint32_t box_c, uint32_t box_w,
                        uint32_t box_h, uint32_t elem_stride) {
  auto enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MGV_ERR_CUDA;
  }
  MGV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p not 16-byte aligned", base);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(C) * W * 2,
                           static_cast<cuuint64_t>(C) * W * H * 2};
  // with an element stride s the box spans box*s tensor elements and loads ceil(span / s) of them
  cuuint32_t box[4] = {box_c, box_w * elem_stride, box_h * elem_stride, 1};
  cuuint32_t estr[4] = {1, elem_stride, elem_stride, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(nhwc C%d W%d H%d N%d box %u,%u,%u stride %u) failed: CUresult %d", C, W, H, N,
              box_c, box_w, box_h, elem_stride, static_cast<int>(r));
    return MGV_ERR_CUDA;
  }
  return MGV_OK;
}

}  // namespace mgv

using namespace mgv;

#define MGV_API_BEGIN try {
#define MGV_API_END                                         \
  }                                                         \
  catch (const std::exception& e) {                         \
    mgv::set_error("C++ exception: %s", e.what());          \
    return MGV_ERR_CUDA;                                    \
  }                                                         \
  catch (...) {                                             \
    mgv::set_error("unknown C++ exception");                \
    return MGV_ERR_CUDA;                                    \
  }

extern "C" {

int mgv_version(void) { return MGV_VERSION; }
const char* mgv_last_error(void) { return mgv::get_error(); }
int mgv_device_check(void) { return mgv::check_device(); }

int mgv_vq_argmin(const float* z, const float* codebook, int B, int C, int HW, int K, int64_t* idx_out, float* dmin_out,
                  mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  return vq_argmin(z, codebook, B, C, HW, K, reinterpret_cast<long long*>(idx_out), dmin_out,
                   static_cast<cudaStream_t>(stream));
  MGV_API_END
}

int mgv_vq_finish(const float* z, const float* codebook, const int64_t* idx, int B, int C, int HW, int K,
                  float commitment_cost, float* quantized, float* encodings, float* loss_out, float* perplexity_out,
                  void* workspace, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  return vq_finish(z, codebook, reinterpret_cast<const long long*>(idx), B, C, HW, K, commitment_cost, quantized,
                   encodings, loss_out, perplexity_out, workspace, static_cast<cudaStream_t>(stream));
  MGV_API_END
}

int mgv_vq_gather(const int64_t* idx, const float* codebook, int64_t n_vec, int C, int HW, int K, float* out,
                  int* bad_index_flag, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  return vq_gather(reinterpret_cast<const long long*>(idx), codebook, n_vec, C, HW, K, out, bad_index_flag,
                   static_cast<cudaStream_t>(stream));
  MGV_API_END
}

int mgv_gpt_create(const mgv_gpt_config* cfg, mgv_gpt_t** out) {
  MGV_API_BEGIN
  static_assert(sizeof(mgv_gpt_config) == sizeof(GptConfig), "config layout");
  return gpt_create(reinterpret_cast<const GptConfig*>(cfg), reinterpret_cast<Gpt**>(out));
  MGV_API_END
}
int mgv_gpt_destroy(mgv_gpt_t* g) {
  MGV_API_BEGIN
  return gpt_destroy(reinterpret_cast<Gpt*>(g));
  MGV_API_END
}
int mgv_gpt_load_weight(mgv_gpt_t* g, const char* name, const float* src, int64_t numel, mgv_stream_t stream) {
  MGV_API_BEGIN
  return gpt_load_weight(reinterpret_cast<Gpt*>(g), name, src, numel, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_gpt_forward(mgv_gpt_t* g, const int64_t* idx, int B, int t, const float* prefix_emb, const int64_t* cls, int m,
                    float* logits_out, float* att_out, mgv_stream_t stream) {
  MGV_API_BEGIN
  return gpt_forward(reinterpret_cast<Gpt*>(g), reinterpret_cast<const long long*>(idx), B, t, prefix_emb,
                     reinterpret_cast<const long long*>(cls), m, logits_out, att_out, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_gpt_generate(mgv_gpt_t* g, const int64_t* x0, int B, int t0, const float* prefix_emb, const int64_t* cls, int m,
                     int steps, float temperature, int do_sample, int top_k, uint64_t seed, int64_t* x_out,
                     float* att_out, int use_graph, mgv_stream_t stream) {
  MGV_API_BEGIN
  return gpt_generate(reinterpret_cast<Gpt*>(g), reinterpret_cast<const long long*>(x0), B, t0, prefix_emb,
                      reinterpret_cast<const long long*>(cls), m, steps, temperature, do_sample, top_k, seed,
                      reinterpret_cast<long long*>(x_out), att_out, use_graph, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_gpt_cross_entropy(mgv_gpt_t* g, const float* logits, const int64_t* targets, int64_t rows, int V, float* loss_out,
                          mgv_stream_t stream) {
  MGV_API_BEGIN
  return gpt_cross_entropy(reinterpret_cast<Gpt*>(g), logits, reinterpret_cast<const long long*>(targets), rows, V, loss_out,
                           static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int64_t mgv_gpt_last_launches(const mgv_gpt_t* g) { return gpt_last_launches(reinterpret_cast<const Gpt*>(g)); }
// ---- training step (gpt_train.cu)
int64_t mgv_gpt_train_numel(mgv_gpt_t* g) { return gpt_train_numel(reinterpret_cast<Gpt*>(g)); }
int mgv_gpt_train_layout(mgv_gpt_t* g, const char* name, int64_t* offset, int64_t* numel, int* decay) {
  MGV_API_BEGIN
  long long o = 0, n = 0;
  const int rc = gpt_train_layout(reinterpret_cast<Gpt*>(g), name, &o, &n, decay);
  if (rc == MGV_OK) { *offset = o; *numel = n; }
  return rc;
  MGV_API_END
}
int mgv_gpt_train_bind(mgv_gpt_t* g, float* flat_params, float* flat_grads, mgv_stream_t stream) {
  MGV_API_BEGIN
  return gpt_train_bind(reinterpret_cast<Gpt*>(g), flat_params, flat_grads, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_gpt_train_forward(mgv_gpt_t* g, const int64_t* idx, int B, int t, const int64_t* cls, int m, const int64_t* targets,
                          float p_embd, float p_resid, float p_attn, uint64_t seed, float* loss_out, mgv_stream_t stream) {
  MGV_API_BEGIN
  return gpt_train_forward(reinterpret_cast<Gpt*>(g), reinterpret_cast<const long long*>(idx), B, t,
                           reinterpret_cast<const long long*>(cls), m, reinterpret_cast<const long long*>(targets), p_embd,
                           p_resid, p_attn, seed, loss_out, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_gpt_train_backward(mgv_gpt_t* g, int layer_hi, int layer_lo, mgv_stream_t stream) {
  MGV_API_BEGIN
  return gpt_train_backward(reinterpret_cast<Gpt*>(g), layer_hi, layer_lo, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_gpt_train_adamw(mgv_gpt_t* g, float* m, float* v, float lr, float beta1, float beta2, float eps, float weight_decay,
                        int64_t step, float grad_scale, mgv_stream_t stream) {
  MGV_API_BEGIN
  return gpt_train_adamw(reinterpret_cast<Gpt*>(g), m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                         static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_test_dropout_mask(uint64_t seed, unsigned stream_id, float p, int64_t n, unsigned char* out, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  return train_drop_mask(make_drop(p, seed), stream_id, n, out, static_cast<cudaStream_t>(stream));
  MGV_API_END
}

int mgv_gpt_set_deterministic(mgv_gpt_t* g, int on) {
  MGV_API_BEGIN
  return gpt_set_deterministic(reinterpret_cast<Gpt*>(g), on);
  MGV_API_END
}
int mgv_gpt_set_step_logits(mgv_gpt_t* g, float* buf) {
  MGV_API_BEGIN
  return gpt_set_step_logits(reinterpret_cast<Gpt*>(g), buf);
  MGV_API_END
}

int mgv_vqvae_create(int num_embeddings, int embedding_dim, mgv_vqvae_t** out) {
  MGV_API_BEGIN
  return vqvae_create(num_embeddings, embedding_dim, reinterpret_cast<Vqvae**>(out));
  MGV_API_END
}
int mgv_vqvae_destroy(mgv_vqvae_t* v) {
  MGV_API_BEGIN
  return vqvae_destroy(reinterpret_cast<Vqvae*>(v));
  MGV_API_END
}
int mgv_vqvae_load_weight(mgv_vqvae_t* v, const char* name, const float* src, int64_t numel, mgv_stream_t stream) {
  MGV_API_BEGIN
  return vqvae_load_weight(reinterpret_cast<Vqvae*>(v), name, src, numel, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_vqvae_decode_codes(mgv_vqvae_t* v, const int64_t* idx, int B, float* mel_out, mgv_stream_t stream) {
  MGV_API_BEGIN
  return vqvae_decode(reinterpret_cast<Vqvae*>(v), reinterpret_cast<const long long*>(idx), nullptr, B, mel_out,
                      static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_vqvae_decode(mgv_vqvae_t* v, const float* quant_bchw, int B, float* mel_out, mgv_stream_t stream) {
  MGV_API_BEGIN
  return vqvae_decode(reinterpret_cast<Vqvae*>(v), nullptr, quant_bchw, B, mel_out, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_vqvae_encode(mgv_vqvae_t* v, const float* mel, int B, float* z_out, mgv_stream_t stream) {
  MGV_API_BEGIN
  return vqvae_encode(reinterpret_cast<Vqvae*>(v), mel, B, z_out, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int64_t mgv_vqvae_last_launches(const mgv_vqvae_t* v) { return vqvae_last_launches(reinterpret_cast<const Vqvae*>(v)); }

int mgv_test_gemm(int impl, const void* A, const void* B, int M, int N, int K, int epi, const float* bias, void* out,
                  const void* resid, int bn, int split_k, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  GemmArgs a;
  a.A = A; a.B = B; a.M = M; a.N = N; a.K = K;
  a.epi = epi; a.bias = bias; a.out = out; a.resid = resid;
  a.bn = bn; a.split_k = split_k;
  a.stream = static_cast<cudaStream_t>(stream);
  return impl == 0 ? gemm_bf16_tc(a) : gemm_bf16_ref(a);
  MGV_API_END
}

int mgv_test_gemm_swapab(int impl, const void* W, const void* X, int M, int N, int K, int epi, const float* bias,
                         void* out, const void* resid, int bn, int split_k, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  GemmArgs a;
  a.A = W; a.B = X; a.M = M; a.N = N; a.K = K;
  a.epi = epi; a.bias = bias; a.out = out; a.resid = resid;
  a.bn = bn; a.split_k = split_k;
  a.transpose_out = true;
  a.stream = static_cast<cudaStream_t>(stream);
  return impl == 0 ? gemm_bf16_tc(a) : gemm_bf16_ref(a);
  MGV_API_END
}

int mgv_test_gemm_fold(int mode, const void* W, const float* src, int Nw, int B, int K, const float* gamma_or_sw,
                       const float* beta_or_bp, const float* bias, const float* stats_in, int nparts, int ln_dim,
                       float* out, int kbps, int staging_warps, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (mode == FOLD_GELU) {
    LnFold in;
    in.stats = reinterpret_cast<const float2*>(stats_in); in.nparts = nparts; in.stride = B; in.dim = ln_dim;
    in.sw = gamma_or_sw; in.bp = beta_or_bp;
    return gemm_decode_fold(FOLD_GELU, W, Nw, K, src, B, nullptr, nullptr, 0, &in, bias, out, Nw, kbps, staging_warps % 100, staging_warps >= 100 ? 64 : 32, false, s);
  }
  // FOLD_LN: raw accumulation + statistics, then the consumer-side correction (what attention / FC2 / the sampler do)
  const int parts = ceil_div(K / 64, kbps);
  float *sw = nullptr, *bp = nullptr;
  float2* stats = nullptr;
  MGV_CHECK_CUDA(cudaMalloc(&sw, static_cast<size_t>(Nw) * 4));
  MGV_CHECK_CUDA(cudaMalloc(&bp, static_cast<size_t>(Nw) * 4));
  MGV_CHECK_CUDA(cudaMalloc(&stats, static_cast<size_t>(parts) * B * sizeof(float2)));
  int rc = gpt_fold_prepare(W, Nw, K, gamma_or_sw, beta_or_bp, bias, sw, bp, s);
  if (rc == MGV_OK)
    rc = gemm_decode_fold(FOLD_LN, W, Nw, K, src, B, gamma_or_sw, stats, B, nullptr, nullptr, out, Nw, kbps, staging_warps % 100, staging_warps >= 100 ? 64 : 32, false, s);
  if (rc == MGV_OK) {
    LnFold f;
    f.stats = stats; f.nparts = parts; f.stride = B; f.dim = K; f.sw = sw; f.bp = bp;
    rc = gpt_fold_apply(out, B, Nw, f, s);
  }
  cudaStreamSynchronize(s);
  cudaFree(sw); cudaFree(bp); cudaFree(stats);
  return rc;
  MGV_API_END
}

int mgv_test_conv3x3(int impl, const void* x, const void* w, const float* bias, int n_img, int Hin, int Win, int Cin,
                     int Cout, int stride, void* out, const void* resid, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  GemmArgs a;
  a.a_mode = A_CONV3x3;
  a.A = x; a.B = w;
  a.n_img = n_img; a.Hin = Hin; a.Win = Win; a.Cin = Cin;
  a.stride = stride;
  if (stride == 1) {
    a.H = Hin; a.W = Win; a.pad = 1;
  } else {
    // Downsample: F.pad(x, (0,1,0,1)) then conv3x3 stride 2 pad 0 (big_model_attn_gan.py:151-159)
    a.H = (Hin + 1 - 3) / 2 + 1; a.W = (Win + 1 - 3) / 2 + 1; a.pad = 0;
  }
  a.M = n_img * a.H * a.W; a.N = Cout; a.K = 9 * Cin;
  a.epi = resid ? EPI_BF16_RESID : EPI_BF16;
  a.bias = bias; a.out = out; a.resid = resid;
  a.bn = (Cout % 128 == 0) ? 128 : (Cout % 64 == 0 ? 64 : 32);
  a.stream = static_cast<cudaStream_t>(stream);
  return impl == 0 ? gemm_bf16_tc(a) : gemm_bf16_ref(a);
  MGV_API_END
}

int mgv_test_attention_prefill(int impl, const void* qkv, int B, int T, int nh, void* y, int64_t* trace, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  MGV_REQUIRE(qkv && y && B >= 0 && nh >= 1, "attention_prefill: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (impl == 0) {
    MGV_REQUIRE(gpt_attention_prefill_tc_supported(T), "attention_prefill: T=%d is not covered by the tcgen05 kernel", T);
    return gpt_attention_prefill_tc(static_cast<const __nv_bfloat16*>(qkv), B, T, nh, static_cast<__nv_bfloat16*>(y), s,
                                    reinterpret_cast<long long*>(trace));
  }
  return gpt_attention_prefill_mma(static_cast<const __nv_bfloat16*>(qkv), B, T, nh, 0, static_cast<__nv_bfloat16*>(y), nullptr, 0,
                                   nullptr, nullptr, 0, s);
  MGV_API_END
}

int mgv_test_conv_upsample(int impl, const void* x, const float* w_oihw, const float* bias, int n_img, int H, int W, int Cin,
                           int Cout, void* out, void* scratch, mgv_stream_t stream) {
  MGV_API_BEGIN
  MGV_TRY(check_device());
  MGV_REQUIRE(x && w_oihw && out && scratch, "conv_upsample: null");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* wp = static_cast<__nv_bfloat16*>(scratch);
  MGV_TRY(vqvae_upsample_phase_weights(w_oihw, Cout, Cin, wp, s));
  for (int ph = 0; ph < 4; ++ph) {
    GemmArgs a;
    a.a_mode = A_CONV3x3;
    a.A = x; a.B = wp + static_cast<size_t>(ph) * Cout * 4 * Cin;
    a.n_img = n_img; a.Hin = H; a.Win = W; a.Cin = Cin; a.stride = 1;
    a.H = H; a.W = W;
    a.taps_x = 2; a.pad = 1 - (ph & 1); a.pad_y = 1 - (ph >> 1);
    a.out_scale = 2; a.out_oy = ph >> 1; a.out_ox = ph & 1;
    a.M = n_img * H * W; a.N = Cout; a.K = 4 * Cin;
    a.epi = EPI_BF16;
    a.bias = bias; a.out = out;
    a.bn = (Cout % 128 == 0) ? 128 : (Cout % 64 == 0 ? 64 : 32);
    a.stream = s;
    MGV_TRY(impl == 0 ? gemm_bf16_tc(a) : gemm_bf16_ref(a));
  }
  return MGV_OK;
  MGV_API_END
}

int mgv_melgan_create(int n_mel_channels, int ngf, int n_residual_layers, mgv_melgan_t** out) {
  MGV_API_BEGIN
  return melgan_create(reinterpret_cast<Melgan**>(out), n_mel_channels, ngf, n_residual_layers);
  MGV_API_END
}
int mgv_melgan_destroy(mgv_melgan_t* m) {
  MGV_API_BEGIN
  return melgan_destroy(reinterpret_cast<Melgan*>(m));
  MGV_API_END
}
int mgv_melgan_load_weight(mgv_melgan_t* m, const char* name, const float* src, int64_t numel, mgv_stream_t stream) {
  MGV_API_BEGIN
  return melgan_load_weight(reinterpret_cast<Melgan*>(m), name, src, numel, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_melgan_reset_biases(mgv_melgan_t* m, mgv_stream_t stream) {
  MGV_API_BEGIN
  return melgan_reset_biases(reinterpret_cast<Melgan*>(m), static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int mgv_melgan_forward(mgv_melgan_t* m, const float* mel, int B, int T, float* wave_out, mgv_stream_t stream) {
  MGV_API_BEGIN
  return melgan_forward(reinterpret_cast<Melgan*>(m), mel, B, T, wave_out, static_cast<cudaStream_t>(stream));
  MGV_API_END
}
int64_t mgv_melgan_last_launches(const mgv_melgan_t* m) { return melgan_last_launches(reinterpret_cast<const Melgan*>(m)); }

}  // extern "C"
