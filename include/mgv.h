/* libmgv -- C ABI of the B200-native token hot path of karchkha/MelSpec_GPT_VQVAE.
 *
 * The reference has no FFI of its own: its boundary for this path is the Python
 * nn.Module surface.  The drop-in Python modules in
 * melspec_gpt_vqvae_b200/ keep that surface and call the entry points below through
 * ctypes; each entry point cites the reference method it replaces (paths relative to the
 * reference repo root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless stated otherwise;
 *     tensors are dense, row-major ("contiguous" in torch terms);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls only
 *     enqueue work unless documented as synchronising;
 *   - return value 0 = success, nonzero = error code (MGV_ERR_*); the message is available
 *     from mgv_last_error() on the calling thread.  Nothing throws or aborts across the ABI;
 *   - there is no CPU fallback: a missing / non-sm_100 device is MGV_ERR_DEVICE;
 *   - opaque handles (mgv_gpt_t, mgv_vqvae_t) own packed bf16 weight copies, KV cache and
 *     workspaces; they are bound to the device that was current at creation and are not
 *     thread-safe per handle.
 */
#ifndef MGV_H_
#define MGV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGV_VERSION 100 /* 0.1.0 */

enum {
  MGV_OK = 0,
  MGV_ERR_INVALID = 1,
  MGV_ERR_CUDA = 2,
  MGV_ERR_DEVICE = 3,
  MGV_ERR_STATE = 4,
  MGV_ERR_UNSUPPORTED = 5
};

typedef void* mgv_stream_t;

int mgv_version(void);
const char* mgv_last_error(void);
/* MGV_OK iff the current CUDA device is compute capability 10.x (B200). */
int mgv_device_check(void);

/* ------------------------------------------------------------------ (1) quantiser ---- */

/* VectorQuantizer.forward, index part  (vqvae/big_model_attn_gan.py:19-33).
 * z_bchw: fp32 (B, C, H*W) -- the encoder output as is (no BHWC permute needed);
 * codebook: fp32 (K, C) = _embedding.weight;  C in {64,128,192,256};  K up to 65536: codebooks beyond 128 codes
 *   (e.g. the reference's 1024-code VGGSound variant) are searched in passes of 128 codes, lowest index still wins;
 * idx_out: int64 (B*H*W) in (b, h, w) order (= encoding_indices.squeeze(1));
 * dmin_out: fp32 (B*H*W), the winning distance; optional for num_embeddings <= 128, REQUIRED beyond (the 128-code
 * passes hand the running minimum to each other through it).
 * Distances are fp32, d = (|x|^2 + |e|^2) - 2<x,e>, every sum a sequential fmaf chain over
 * the channel index; first index wins ties (torch.argmin).  oracle/vq_oracle.c restates it.
 * Implementation: a tensor-core (TF32) prefilter narrows every vector to the codes within a proven error margin of the
 * minimum, the exact chain is evaluated for those only; results are bit-identical to evaluating all K distances.  With
 * dmin_out == NULL (and K <= 128) vectors with a single candidate skip the exact evaluation.  B*H*W must be below 2^31 - 1024;
 * the call allocates 16 bytes per vector of stream-ordered scratch (cudaMallocFromPoolAsync on `stream`). */
int mgv_vq_argmin(const float* z_bchw, const float* codebook, int B, int C, int HW, int K,
                  int64_t* idx_out, float* dmin_out, mgv_stream_t stream);

/* Rest of VectorQuantizer.forward (:36-54) given the indices: straight-through quantized
 * (B, C, H*W) BCHW, one-hot encodings (B*H*W, K), loss and perplexity scalars (device
 * pointers to 1 float).  Any output pointer may be NULL.  workspace: >= 8 + 4*K bytes,
 * 8-byte aligned. */
int mgv_vq_finish(const float* z_bchw, const float* codebook, const int64_t* idx, int B, int C, int HW,
                  int K, float commitment_cost, float* quantized_bchw, float* encodings, float* loss_out,
                  float* perplexity_out, void* workspace, mgv_stream_t stream);

/* VectorQuantizer.get_codebook_entry (:56-71): out[b, c, p] = codebook[idx[b*HW + p], c]
 * (HW > 0, shape given) or out[n, c] = codebook[idx[n], c] (HW == 0, shape None).
 * bad_index_flag: optional device int set to 1 if any index is outside [0, K). */
int mgv_vq_gather(const int64_t* idx, const float* codebook, int64_t n_vec, int C, int HW, int K,
                  float* out, int* bad_index_flag, mgv_stream_t stream);

/* ------------------------------------------------------------------ (2) minGPT ------- */

typedef struct mgv_gpt mgv_gpt_t;

typedef struct {
  int vocab_size;  /* config/config_GPT_vas.py: 128 */
  int block_size;  /* 266 */
  int n_layer;     /* 24 */
  int n_head;      /* 16 */
  int n_embd;      /* 1024 (n_embd / n_head must be 64) */
  int class_size;  /* 8; 0 = plain GPT without `embedder` */
  int n_unmasked;  /* transformer/minGPT.py:67-68 */
  int head_out;    /* 0 = vocab_size; else GPT(last_linear=...) (:144-149) */
} mgv_gpt_config;

/* GPT.__init__ / GPTClass.__init__ (transformer/minGPT.py:123-154, :205-207). */
int mgv_gpt_create(const mgv_gpt_config* cfg, mgv_gpt_t** out);
int mgv_gpt_destroy(mgv_gpt_t* g);

/* Pack one state_dict tensor (fp32, device) into the handle; `name` is the reference
 * state_dict key relative to the GPT module ("tok_emb.weight", "pos_emb",
 * "blocks.3.attn.query.weight", "blocks.3.mlp.0.bias", "ln_f.weight", "head.weight",
 * "embedder.weight", ...).  "blocks.N.attn.mask" is accepted and ignored.  Linear weights
 * are converted to bf16 (query/key/value fused into one 3C x C matrix). */
int mgv_gpt_load_weight(mgv_gpt_t* g, const char* name, const float* src, int64_t numel, mgv_stream_t stream);

/* GPT.forward(idx, embeddings) / GPTClass.forward(idx, token) (:168-199, :209-212),
 * eval mode (dropout off).  Sequence = m prefix rows + t token rows.
 *   idx: int64 (B, t);  prefix_emb: fp32 (B, m, C) or NULL;  cls: int64 (B) class ids used
 *   when prefix_emb is NULL and m == 1 (GPTClass);  logits_out: fp32 (B, m+t, head_out);
 *   att_out: optional fp32 (B, n_head, m+t, m+t) = last block's attention probabilities.
 * Synchronises `stream` (reports out-of-range token / class ids as MGV_ERR_INVALID). */
int mgv_gpt_forward(mgv_gpt_t* g, const int64_t* idx, int B, int t, const float* prefix_emb, const int64_t* cls,
                    int m, float* logits_out, float* att_out, mgv_stream_t stream);

/* Lit_minGPT.sample (:293-360) with a KV cache and a CUDA-graph decode loop.
 *   x0: int64 (B, t0) prompt (t0 may be 0);  x_out: int64 (B, t0+steps), first t0 columns = x0;
 *   att_out: optional fp32 (B, n_head, Tf, Tf), Tf = m + t0 + steps - 1 (the attention the
 *   reference returns from its last forward);  do_sample 0 = greedy (torch.topk(probs,1)),
 *   1 = multinomial;  top_k 0 = None;  seed keys the Philox stream (row b, position p).
 * Fails with MGV_ERR_INVALID when t0 + steps - 1 + m > block_size (the reference's assert).
 * Synchronises. */
int mgv_gpt_generate(mgv_gpt_t* g, const int64_t* x0, int B, int t0, const float* prefix_emb, const int64_t* cls,
                     int m, int steps, float temperature, int do_sample, int top_k, uint64_t seed,
                     int64_t* x_out, float* att_out, int use_graph, mgv_stream_t stream);

/* Per-row cross entropy, loss_out[r] = logsumexp(logits[r, :V]) - logits[r, targets[r]]:
 * F.cross_entropy in GPT.forward(targets=...) (transformer/minGPT.py:195-197, mean taken by the caller) and
 * nn.CrossEntropyLoss(weight=ones, reduction='none') in GPTDecoder.reconstruct_error (transformer/decoders.py:21,64-68).
 *   logits: fp32 (rows, V) contiguous;  targets: int64 (rows,);  loss_out: fp32 (rows,).
 * A target outside [0, V) fails with MGV_ERR_INVALID (torch raises an index error).  Synchronises. */
int mgv_gpt_cross_entropy(mgv_gpt_t* g, const float* logits, const int64_t* targets, int64_t rows, int V, float* loss_out,
                          mgv_stream_t stream);

/* number of kernels libmgv launched in the last forward / generate call on this handle */
int64_t mgv_gpt_last_launches(const mgv_gpt_t* g);

/* ---- minGPT training step (BASELINE config 4).  reference: Lit_minGPT.training_step / shared_step
 * transformer/minGPT.py:413-422 (teacher-forced logits + mean cross entropy), GPT.forward :168-199 with its three dropouts,
 * configure_optimizers :618-665 (AdamW, betas (0.9, 0.95), weight decay 0.01 on Linear weights only),
 * DDP gradient averaging GPT_VAE_train.py:172-174 (done by the host with NCCL between mgv_gpt_train_backward calls).
 *
 * The caller owns two flat fp32 device buffers of mgv_gpt_train_numel() elements -- master parameters and gradients -- whose
 * layout (one 256-byte aligned segment per state_dict tensor, blocks contiguous and in order, q|k|v adjacent) is reported by
 * mgv_gpt_train_layout.  The handle keeps bf16 / transposed copies for the tcgen05 GEMMs and refreshes them in the optimizer. */
int64_t mgv_gpt_train_numel(mgv_gpt_t* g);
/* offset / numel (elements) of state_dict tensor `name` in the flat buffers; *decay = 1 if weight decay applies to it */
int mgv_gpt_train_layout(mgv_gpt_t* g, const char* name, int64_t* offset, int64_t* numel, int* decay);
/* attach the flat buffers (the handle must already hold the same weights through mgv_gpt_load_weight) */
int mgv_gpt_train_bind(mgv_gpt_t* g, float* flat_params, float* flat_grads, mgv_stream_t stream);
/* forward with dropout (probabilities in [0,1), masks from a counter hash of `seed`): idx (B, t) int64, cls (B) int64 with
 * m = 1 (or NULL, m = 0), targets (B, m + t) int64; *loss_out (device fp32) = mean cross entropy.  Saves activations. */
int mgv_gpt_train_forward(mgv_gpt_t* g, const int64_t* idx, int B, int t, const int64_t* cls, int m, const int64_t* targets,
                          float p_embd, float p_resid, float p_attn, uint64_t seed, float* loss_out, mgv_stream_t stream);
/* backward of the last forward through blocks layer_hi-1 .. layer_lo.  layer_hi == n_layer first clears the gradient buffer
 * and runs the head / ln_f part; layer_lo == 0 also runs the embedding backward.  Calling it bucket by bucket lets the host
 * all-reduce the finished (contiguous) slice of the gradient buffer while the next bucket is computed. */
int mgv_gpt_train_backward(mgv_gpt_t* g, int layer_hi, int layer_lo, mgv_stream_t stream);
/* torch.optim.AdamW step on the flat buffers (m, v: fp32 state, same length); gradients are multiplied by grad_scale first */
int mgv_gpt_train_adamw(mgv_gpt_t* g, float* m, float* v, float lr, float beta1, float beta2, float eps, float weight_decay,
                        int64_t step, float grad_scale, mgv_stream_t stream);
/* tests: keep flags (1 / 0) of elements 0..n-1 of dropout stream `stream_id` (1: embedding; 16 + 4l: attention of block l;
 * 17 + 4l / 18 + 4l: residual dropout after proj / mlp), element index = row-major index of the dropped tensor */
int mgv_test_dropout_mask(uint64_t seed, unsigned stream_id, float p, int64_t n, unsigned char* out, mgv_stream_t stream);

/* Deterministic decode switch.  The default decode loop reduces split-K partial sums with fp32 atomics whose order depends
 * on CTA timing, so sampled tokens can differ run to run at near-ties (the reference's greedy path is deterministic).
 * on != 0 selects a schedule in which every decode GEMM owns its full K: bit-reproducible, about 2x slower. */
int mgv_gpt_set_deterministic(mgv_gpt_t* g, int on);

/* One-shot inspection hook for parity tests: the NEXT mgv_gpt_generate call on this handle also writes the logits of
 * every decode step (after the temperature division, before top-k; reference transformer/minGPT.py:346) to
 * buf (steps, B, V) fp32, device memory owned by the caller.  NULL cancels a pending request. */
int mgv_gpt_set_step_logits(mgv_gpt_t* g, float* buf);

/* ------------------------------------------------------------------ (3) VQVAE -------- */

typedef struct mgv_vqvae mgv_vqvae_t;

/* LitVQVAE(num_embeddings, embedding_dim) with the module-level architecture constants of
 * vqvae/big_model_attn_gan.py:521-530 (ch 128, ch_mult 1,1,2,2,4, 2 res blocks, attention
 * at the 5x53 level, z_channels 256, 80x848 mels). */
int mgv_vqvae_create(int num_embeddings, int embedding_dim, mgv_vqvae_t** out);
int mgv_vqvae_destroy(mgv_vqvae_t* v);

/* state_dict key relative to LitVQVAE ("_decoder.up.0.block.1.conv1.weight",
 * "post_quant_conv.bias", "_vq_vae._embedding.weight", "_encoder.conv_in.weight", ...);
 * "discriminator.*" keys are accepted and ignored.  src: fp32 device. */
int mgv_vqvae_load_weight(mgv_vqvae_t* v, const char* name, const float* src, int64_t numel, mgv_stream_t stream);

/* Lit_minGPT.decode_to_img after code_reader (transformer/minGPT.py:515-528):
 * get_codebook_entry + post_quant_conv + Decoder.forward (big_model_attn_gan.py:56-71,
 * :610-614, :361-392).  idx: int64 (B, H*W) row-major code grid; mel_out: fp32 (B,1,80,848).
 * Returns after `stream` has completed (the out-of-range index flag is read back).  From the second
 * call at a batch size on, the ~150 launches are captured into / replayed from a CUDA graph owned by
 * the handle: idx and mel_out are copied through handle-owned staging buffers on `stream`, so the
 * call stays stream-ordered and the caller's pointers may change from call to call; weights loaded
 * later are seen by the graph (they are rewritten in place).  One call at a time per handle.
 * Upsample (:182-186) is evaluated as four 2x2 convolutions over the low-res tensor (same result up
 * to the bf16 rounding of the pre-summed weights). */
int mgv_vqvae_decode_codes(mgv_vqvae_t* v, const int64_t* idx, int B, float* mel_out, mgv_stream_t stream);

/* LitVQVAE.decode(quant) (:610-614): quant fp32 (B, 256, 5, 53) BCHW -> mel (B,1,80,848). */
int mgv_vqvae_decode(mgv_vqvae_t* v, const float* quant_bchw, int B, float* mel_out, mgv_stream_t stream);

/* LitVQVAE.encode(x) (:604-608): mel fp32 (B,1,80,848) -> z fp32 (B, 256, 5, 53) BCHW. */
int mgv_vqvae_encode(mgv_vqvae_t* v, const float* mel, int B, float* z_out, mgv_stream_t stream);

int64_t mgv_vqvae_last_launches(const mgv_vqvae_t* v);

/* ------------------------------------------------------------------ test hooks ------- */

/* D[M,N] = A[M,K] B[N,K]^T (+bias) through the tcgen05 GEMM (impl 0) or the SIMT fp32
 * reference kernel (impl 1).  A, B bf16; epi as GemmEpilogue in csrc/gemm_tc.cuh.
 * Used by tests/ to cross-check the tensor-core path on device. */
int mgv_test_gemm(int impl, const void* A, const void* B, int M, int N, int K, int epi, const float* bias,
                  void* out, const void* resid, int bn, int split_k, mgv_stream_t stream);

/* Swap-AB form used by the decode step: W bf16 (M = out features, K) is the 128-row MMA operand, X bf16
 * (N = batch rows, K) the MMA N dimension; out[n * M + m] (+ bias[m]).  N need not be a multiple of 32. */
int mgv_test_gemm_swapab(int impl, const void* W, const void* X, int M, int N, int K, int epi, const float* bias,
                         void* out, const void* resid, int bn, int split_k, mgv_stream_t stream);

/* Decode GEMMs that absorb the LayerNorm / GELU in front of them (gemm_decode_fold.cu; reference math:
 * transformer/minGPT.py:100-119, 186-188).  W bf16 (Nw, K); src fp32 (B, K); out fp32 (B, Nw), accumulated in place.
 * mode 0: out += LayerNorm(src; gamma, beta, eps 1e-5) W^T + bias, computed as the folded form the decode chain uses
 *         (raw W bf16(gamma.src) accumulation + per-K-slice row statistics, then rstd * (acc - mu * W gamma) + W beta + bias);
 *         out must hold zeros.  gamma_or_sw = gamma (K), beta_or_bp = beta (K).
 * mode 1: out += gelu_erf(rstd_b * (src - mu_b * sw) + bp) W^T + bias with mu / rstd from stats_in (nparts, B, 2) =
 *         partial (sum, sum of squares) over ln_dim elements.  gamma_or_sw = sw (K), beta_or_bp = bp (K).
 * kbps = 64-wide k-blocks per CTA (1..4); staging_warps = 4 or 8, plus 100 to select 64 sequences per CTA (default 32). */
int mgv_test_gemm_fold(int mode, const void* W, const float* src, int Nw, int B, int K, const float* gamma_or_sw,
                       const float* beta_or_bp, const float* bias, const float* stats_in, int nparts, int ln_dim,
                       float* out, int kbps, int staging_warps, mgv_stream_t stream);

/* 3x3 convolution (stride 1 pad 1, or stride 2 with the reference's (0,1,0,1) padding) over
 * NHWC bf16 input through the implicit-GEMM path (impl 0) or the SIMT reference (impl 1).
 * w: bf16 (Cout, 3, 3, Cin). */
int mgv_test_conv3x3(int impl, const void* x_nhwc, const void* w, const float* bias, int n_img, int Hin, int Win,
                     int Cin, int Cout, int stride, void* out_nhwc, const void* resid_nhwc, mgv_stream_t stream);

/* Causal self-attention of a whole sequence (CausalSelfAttention.forward transformer/minGPT.py:70-92, head dim 64):
 * qkv bf16 [B*T, 3*nh*64] = [q | k | v] -> y bf16 [B*T, nh*64].  impl 0 = tcgen05 kernel, impl 1 = mma.sync kernel.
 * trace (impl 0, optional): 48 int64 clock stamps of one CTA. */
int mgv_test_attention_prefill(int impl, const void* qkv, int B, int T, int nh, void* y, int64_t* trace, mgv_stream_t stream);

/* Upsample.forward (vqvae/big_model_attn_gan.py:182-186: nearest 2x, then 3x3 conv pad 1) in the phase form the decoder
 * uses: four 2x2 convolutions over the low-res NHWC bf16 input with weights pre-summed from w (fp32 OIHW (Cout, Cin, 3, 3)).
 * impl 0 = tcgen05 implicit GEMM, impl 1 = SIMT reference of the same contract.  out: bf16 (n_img, 2H, 2W, Cout);
 * scratch: bf16 [16 * Cout * Cin] for the phase weights. */
int mgv_test_conv_upsample(int impl, const void* x_nhwc, const float* w_oihw, const float* bias, int n_img, int H, int W,
                           int Cin, int Cout, void* out_nhwc, void* scratch, mgv_stream_t stream);

/* ------------------------------------------------------------------ (5) MelGAN vocoder ---- */
/* The step after the path's end: mel -> waveform with the MelGAN Generator the callbacks use for audio logging
 * (vocoder/modules.py:38-80; callers callbacks/GPT_callbacks.py:93-105, callbacks/GPT_VAE_callbacks.py:84-92).  fp32. */
typedef struct mgv_melgan mgv_melgan_t;

/* Generator(input_size = n_mel_channels, ngf, n_residual_layers)  (vocoder/modules.py:39-77; ratios 8, 8, 2, 2) */
int mgv_melgan_create(int n_mel_channels, int ngf, int n_residual_layers, mgv_melgan_t** out);
int mgv_melgan_destroy(mgv_melgan_t* m);

/* state_dict key of the reference Generator with the weight norm applied by the caller (w = g * v / |v|, the norm
 * over all dimensions but the first: torch.nn.utils.weight_norm, vocoder/modules.py:17-21):
 *   "model.<i>.weight" | "model.<i>.bias" | "model.<i>.block.{2,4}.{weight,bias}" | "model.<i>.shortcut.{weight,bias}".
 * Call mgv_melgan_reset_biases before (re)loading a state_dict: a ResnetBlock's shortcut and last convolution run as
 * one launch whose bias is the sum of both.  src: fp32 device. */
int mgv_melgan_load_weight(mgv_melgan_t* m, const char* name, const float* src, int64_t numel, mgv_stream_t stream);
int mgv_melgan_reset_biases(mgv_melgan_t* m, mgv_stream_t stream);

/* Generator.forward (vocoder/modules.py:79-80): mel fp32 (B, n_mel, T), T >= 4 -> wave fp32 (B, 1, 256 * T) in [-1, 1]. */
int mgv_melgan_forward(mgv_melgan_t* m, const float* mel, int B, int T, float* wave_out, mgv_stream_t stream);
int64_t mgv_melgan_last_launches(const mgv_melgan_t* m);

#ifdef __cplusplus
}
#endif
#endif /* MGV_H_ */
