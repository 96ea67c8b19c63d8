"""Times mgv_gpt_generate (VAS model, bs=64, 265 tokens) for several sequence-group counts (MGV_DECODE_GROUPS) and
split-K tilings, and reports how many sampled tokens agree with the single-group run (same Philox seed).
Diagnostic tool (run under gpurun)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT

dev = torch.device("cuda", 0)
cfg = synthetic.GPT_VAS
sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=False)
B = int(os.environ.get("TUNE_B", "64"))
c = torch.randint(0, 8, (B, 1), generator=torch.Generator().manual_seed(0)).to(dev)
x0 = torch.zeros(B, 0, dtype=torch.long, device=dev)
ref = None


def run(groups, tiles, steps=265, reps=3):
    global ref
    os.environ["MGV_DECODE_GROUPS"] = str(groups)
    if tiles: os.environ["MGV_DECODE_SPLITS"] = tiles
    else: os.environ.pop("MGV_DECODE_SPLITS", None)
    args = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, reconstruct_spec="", device=dev, **cfg)
    lit = Lit_minGPT(args)
    lit.transformer.load_state_dict(sd, strict=False)
    lit = lit.eval().to(dev)
    lit.return_attention = False
    lit.sample_seed = 1234
    try:
        for _ in range(2):
            out = lit.sample(x0, c, steps=steps, sample=True, top_k=100)[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = lit.sample(x0, c, steps=steps, sample=True, top_k=100)[0]
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if ref is None: ref = out.clone()
        first_div = ((out != ref).float().cumsum(1) == 0).sum(1).float().mean().item()
        print("groups=%d tiles=%-14s: %.1f ms per generation, %.1f us/position, %.0f tok/s ; mean tokens before first divergence from run 0: %.0f/%d"
              % (groups, tiles or "default", ms, ms * 1e3 / steps, B * steps / ms * 1e3, first_div, steps), flush=True)
    except Exception as e:
        print("groups=%d tiles=%s FAILED: %r" % (groups, tiles, e), flush=True)
    del lit
    torch.cuda.empty_cache()


if __name__ == "__main__":
    specs = sys.argv[1:] or ["1:", "2:", "4:", "8:", "2:4,8,2,8", "2:4,8,4,16", "4:4,8,2,8", "4:2,4,1,4", "4:2,4,2,8"]
    for sp in specs:
        gr, tl = sp.split(":")
        run(int(gr), tl)
