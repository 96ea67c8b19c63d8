"""Times mgv_gpt_generate (VAS model, bs=64, 265 tokens) for several decode tilings / PDL settings.
Diagnostic tool (run under gpurun); prints one line per configuration."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT

dev = torch.device("cuda", 0)
cfg = synthetic.GPT_VAS
sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=False)
c = torch.randint(0, 8, (64, 1), generator=torch.Generator().manual_seed(0)).to(dev)
x0 = torch.zeros(64, 0, dtype=torch.long, device=dev)


def run(tiles, pdl, steps=265, reps=3):
    if tiles: os.environ["MGV_DECODE_SPLITS"] = tiles
    else: os.environ.pop("MGV_DECODE_SPLITS", None)
    os.environ["MGV_PDL"] = "1" if pdl else "0"
    args = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, reconstruct_spec="", device=dev, **cfg)
    lit = Lit_minGPT(args)
    lit.transformer.load_state_dict(sd, strict=False)
    lit = lit.eval().to(dev)
    lit.return_attention = False
    try:
        for _ in range(2):
            lit.sample(x0, c, steps=steps, sample=True, top_k=100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lit.sample(x0, c, steps=steps, sample=True, top_k=100)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("tiles=%-28s pdl=%d : %.1f ms per generation, %.1f us/position, %.0f tok/s" % (tiles or "default", pdl, ms, ms * 1e3 / steps, 64 * steps / ms * 1e3), flush=True)
    except Exception as e:
        print("tiles=%s pdl=%d FAILED: %r" % (tiles, pdl, e), flush=True)
    del lit
    torch.cuda.empty_cache()


if __name__ == "__main__":
    configs = sys.argv[1:] or ["", "4,16,4,16", "8,8,4,8", "16,16,8,16", "4,8,2,8", "8,16,4,32"]
    for t in configs:
        for pdl in (0, 1):
            run(t, pdl)
