"""Secondary timings (run under gpurun): teacher-forced forward (prefill), extract_codes path (encode + quantise),
decode_to_img, quantiser alone.  Prints one line each with the roofline fraction the SURVEY section 8(d) names."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT
from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE

dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
TF, HBM = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]


def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cfg = synthetic.GPT_VAS
args = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, reconstruct_spec="", device=dev, **cfg)
lit = Lit_minGPT(args); lit.transformer.load_state_dict(synthetic.synthetic_gpt_state_dict(cfg, perturb=False), strict=False)
lit = lit.eval().to(dev)
x = torch.randint(0, 128, (64, 265)).to(dev); c = torch.randint(0, 8, (64, 1)).to(dev)
ms = timeit(lambda: lit(x, c))
fl = 2 * 302442496 * 64 * 265 + 4 * 265 * 265 * 1024 * 24 * 64
print("teacher-forced forward bs=64 T=265: %.2f ms  %.1f TFLOP/s  (%.3f of sustained bf16 peak)  %.0f tok/s" % (ms, fl / ms / 1e9, fl / ms / 1e9 / TF, 64 * 265 / ms * 1e3))
del lit
vq = LitVQVAE(128, 256); vq.load_state_dict(synthetic.synthetic_vqvae_state_dict(128, 256, perturb=False), strict=False); vq = vq.eval().to(dev)
for B in (64, 256):
    mel = torch.rand(B, 1, 80, 848, device=dev) * 2 - 1
    def enc():
        z = vq.encode(mel)
        return vq._vq_vae.encoding_indices(z)
    ms = timeit(enc, n=3, warm=1)
    print("extract_codes path (encode + argmin) B=%d: %.2f ms  %.0f clips/s  encoder %.1f TFLOP/s (%.3f of peak)" % (B, ms, B / ms * 1e3, 143.0e9 * B / ms / 1e9, 143.0e9 * B / ms / 1e9 / TF))
    z = vq.encode(mel)
    msq = timeit(lambda: vq._vq_vae.encoding_indices(z), n=10)
    print("  quantiser alone N=%d: %.1f us  %.0f GB/s algorithmic (%.3f of HBM peak)" % (B * 265, msq * 1e3, B * 265 * 1032 / msq / 1e6, B * 265 * 1032 / msq / 1e6 / HBM))
    del mel, z
codes = torch.randint(0, 128, (64, 265)).to(dev)
ms = timeit(lambda: vq.decode_codes(codes), n=3, warm=1)
print("decode_to_img B=64: %.2f ms  %.0f clips/s  %.1f TFLOP/s (%.3f of peak)" % (ms, 64 / ms * 1e3, 261.3e9 * 64 / ms / 1e9, 261.3e9 * 64 / ms / 1e9 / TF))
del vq
torch.cuda.empty_cache()

# ---- BASELINE config 1 on the GPU: one clip, 265 tokens (latency-bound: the same 8 stages per block at batch 1)
lit = Lit_minGPT(args); lit.transformer.load_state_dict(synthetic.synthetic_gpt_state_dict(cfg, perturb=False), strict=False)
lit = lit.eval().to(dev); lit.return_attention = False
c1 = torch.tensor([[3]], device=dev); x1 = torch.zeros(1, 0, dtype=torch.long, device=dev)
ms = timeit(lambda: lit.sample(x1, c1, steps=265, sample=True, top_k=100), n=3, warm=2)
print("config 1 (one clip, 265 tokens, multinomial top_k=100): %.1f ms  %.0f tok/s  %.0f us/position" % (ms, 265 / ms * 1e3, ms * 1e3 / 265))
del lit
torch.cuda.empty_cache()

# ---- BASELINE config 5: GPT-VAE (GPT-medium: vocab 1024, 24 layers, 1024-d latent), bs=128: encoder -> z -> decoder loss
from melspec_gpt_vqvae_b200.transformer.Lit_GPT_VAE import GPT_VAE
vargs = argparse.Namespace(embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, fix_var=-1.0, device=dev, kl_start=1.0,
                           vocab_size=1024, block_size=265, n_layer=24, n_head=16, n_embd=1024)
vae = GPT_VAE(vargs).eval().to(dev)
xv = torch.randint(0, 1024, (128, 265), device=dev)
ms = timeit(lambda: vae.loss(xv, 1.0, nsamples=1), n=3, warm=2)
flv = 2 * (2 * 302.4e6 * 128 * 265) + 2 * 4 * 265 * 265 * 1024 * 24 * 128      # two 24-layer GPTs (encoder + decoder)
print("config 5 GPT-VAE loss (encoder + reparameterise + decoder CE) bs=128: %.2f ms  %.0f clips/s  %.1f TFLOP/s (%.3f of peak)" % (ms, 128 / ms * 1e3, flv / ms / 1e9, flv / ms / 1e9 / TF))
z = vae.sample_from_inference(xv[:64], 1)
vae.decoder.return_attention = False      # like bench.py: the (B,16,T,T) attention map is an optional 288 MB by-product
ms = timeit(lambda: vae.decode(z, "beam"), n=2, warm=1)
print("config 5 GPT-VAE decode from latent bs=64 (265 tokens, top_k=100, vocab 1024): %.1f ms  %.0f tok/s" % (ms, 64 * 265 / ms * 1e3))
