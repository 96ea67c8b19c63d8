"""One small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_targets.py
Sizes are tiny on purpose (the sanitizer serialises and instruments every access): 2-layer GPT of the VAS width, a few
decode positions, one VQVAE clip, a short MelGAN clip, one training step.  Run under gpurun."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT
from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE, VectorQuantizer
from melspec_gpt_vqvae_b200.vocoder.modules import Generator

which = set(sys.argv[1:]) or {"vq", "gpt", "vqvae", "melgan", "train"}
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
if "vq" in which:
    vq = VectorQuantizer(128, 256, 0.25).to(dev)
    loss, quant, (perp, enc, idx) = vq((torch.randn(3, 256, 5, 53, generator=g) * 0.2).to(dev))      # ragged tile (795 vectors)
    vq.get_codebook_entry(idx.reshape(-1), (3, 5, 53, 256))
    vq2 = VectorQuantizer(300, 256, 0.25).to(dev)                                                     # multi-pass codebook
    vq2((torch.randn(1, 256, 5, 53, generator=g) * 0.2).to(dev))
    print("vq ok")
cfg = dict(synthetic.GPT_VAS, n_layer=2)
if "gpt" in which or "train" in which:
    args = argparse.Namespace(embd_pdrop=0.1, resid_pdrop=0.1, attn_pdrop=0.1, reconstruct_spec="", device=dev, learning_rate=1e-4, **cfg)
    lit = Lit_minGPT(args)
    lit.transformer.load_state_dict(synthetic.synthetic_gpt_state_dict(cfg, seed=1, perturb=True), strict=False)
    lit = lit.to(dev)
if "gpt" in which:
    lit.eval()
    x = torch.randint(0, 128, (3, 265), generator=g).to(dev)
    c = torch.randint(0, 8, (3, 1), generator=g).to(dev)
    lit(x, c)
    lit.sample(x[:, :5], c, steps=6, sample=True, top_k=100)            # prefill of 5 + 6 fold-chain positions + attention map
    lit.return_attention = False
    lit.sample(torch.zeros(3, 0, dtype=torch.long, device=dev), c, steps=4, sample=False)
    print("gpt ok")
if "train" in which:
    lit.train()
    opt = lit.configure_optimizers()
    batch = {"codes": torch.randint(0, 128, (2, 5, 53), generator=g).to(dev), "target": torch.randint(0, 8, (2,), generator=g).to(dev)}
    loss = lit.training_step(batch, 0)
    opt.step()
    print("train ok, loss %.3f" % float(loss))
if "vqvae" in which:
    m = LitVQVAE(128, 256)
    m.load_state_dict(synthetic.synthetic_vqvae_state_dict(128, 256, perturb=True), strict=False)
    m = m.eval().to(dev)
    m.decode_codes(torch.randint(0, 128, (1, 265), generator=g).to(dev))
    m.encode((torch.rand(1, 1, 80, 848, generator=g) * 2 - 1).to(dev))
    print("vqvae ok")
if "melgan" in which:
    gen = Generator(80, 8, 3)
    gen.load_state_dict(synthetic.synthetic_melgan_state_dict(80, 8, 3, seed=2))
    gen = gen.eval().to(dev)
    gen(torch.rand(2, 80, 21, generator=g).to(dev))
    print("melgan ok")
torch.cuda.synchronize()
print("done")
