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
