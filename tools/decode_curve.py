"""Decode-loop time as a function of context: times mgv_gpt_generate (VAS model, bs=64) for several step counts and
prints the cumulative time and the marginal microseconds per position of each context range.
Diagnostic tool (run under gpurun):  python tools/decode_curve.py [B]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT

dev = torch.device("cuda", 0)
cfg = synthetic.GPT_VAS
sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
args = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, reconstruct_spec="", device=dev, **cfg)
lit = Lit_minGPT(args)
lit.transformer.load_state_dict(sd, strict=False)
lit = lit.eval().to(dev)
lit.return_attention = False
c = torch.randint(0, 8, (B, 1), generator=torch.Generator().manual_seed(0)).to(dev)
x0 = torch.zeros(B, 0, dtype=torch.long, device=dev)


def timed(steps, reps=3):
    for _ in range(2):
        lit.sample(x0, c, steps=steps, sample=True, top_k=100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lit.sample(x0, c, steps=steps, sample=True, top_k=100)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


prev_s, prev_t = 0, 0.0
STEPS = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [8, 40, 72, 136, 200, 265]
for steps in STEPS:
    ms = timed(steps)
    print(os.environ.get("TAG", ""), "B=%d steps=%3d: %8.2f ms total, %7.1f us/position overall, %7.1f us/position over contexts %d..%d"
          % (B, steps, ms, ms * 1e3 / steps, (ms - prev_t) * 1e3 / (steps - prev_s), prev_s, steps), flush=True)
    prev_s, prev_t = steps, ms
