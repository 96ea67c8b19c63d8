"""Small drivers for ncu captures (run under gpurun):  python tools/profile_targets.py decode|vqvae|encode|vq|prefill"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import synthetic

mode = sys.argv[1]
dev = torch.device("cuda", 0)
if mode in ("decode", "prefill"):
    from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT
    cfg = synthetic.GPT_VAS
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=False)
    args = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, reconstruct_spec="", device=dev, **cfg)
    lit = Lit_minGPT(args); lit.transformer.load_state_dict(sd, strict=False); lit = lit.eval().to(dev); lit.return_attention = False
    c = torch.randint(0, 8, (64, 1)).to(dev)
    if mode == "decode":
        steps = int(sys.argv[2]) if len(sys.argv) > 2 else 140
        lit.sample(torch.zeros(64, 0, dtype=torch.long, device=dev), c, steps=steps, sample=True, top_k=100)
    else:
        x = torch.randint(0, 128, (64, 265)).to(dev)
        lit(x, c)
elif mode in ("vqvae", "encode"):
    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    m = LitVQVAE(128, 256); m.load_state_dict(synthetic.synthetic_vqvae_state_dict(128, 256, perturb=False), strict=False); m = m.eval().to(dev)
    if mode == "vqvae":
        m.decode_codes(torch.randint(0, 128, (B, 265)).to(dev))
    else:
        m.encode(torch.rand(B, 1, 80, 848, device=dev) * 2 - 1)
elif mode == "vq":
    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import VectorQuantizer
    vq = VectorQuantizer(128, 256, 0.25).to(dev)
    z = torch.randn(256, 256, 5, 53, device=dev) * 0.2
    for _ in range(3):
        vq(z)
torch.cuda.synchronize()
print("done", mode)
