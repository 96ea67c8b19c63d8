"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported
through ref_shim.py) on seeded synthetic weights and inputs.  Runs only in the build
container (the reference does not travel to the GPU box); the fixtures are committed.

    python tests/golden/make_golden.py

Every fixture stores only reference OUTPUTS plus the seeds/config needed to regenerate the
inputs with melspec_gpt_vqvae_b200.synthetic (torch CPU generators are deterministic).
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from melspec_gpt_vqvae_b200 import synthetic  # noqa: E402

torch.set_grad_enabled(False)


def vq_inputs(case: str):
    """Seeded (z, codebook, K, D) for the named quantiser case -- shared with the tests."""
    g = torch.Generator().manual_seed({"trained": 11, "default_init": 12, "ties": 13, "small": 14}[case])
    if case == "trained":      # realistic scales: z std 0.2, codebook N(0, 0.2)
        z = torch.randn(4, 256, 5, 53, generator=g) * 0.2
        cb = torch.randn(128, 256, generator=g) * 0.2
    elif case == "default_init":  # reference init U(+-1/K): all codes nearly equidistant
        z = torch.randn(4, 256, 5, 53, generator=g) * 0.2
        cb = (torch.rand(128, 256, generator=g) * 2 - 1) / 128
    elif case == "ties":       # duplicated codes, z equal to codes, zero vectors
        cb = torch.randn(128, 256, generator=g) * 0.2
        cb[64:] = cb[:64]                      # every code has an exact duplicate
        z = torch.randn(2, 256, 5, 53, generator=g) * 0.2
        zf = z.permute(0, 2, 3, 1).reshape(-1, 256)
        zf[:100] = cb[torch.arange(100) % 128]  # exact hits
        zf[100:110] = 0.0
        z = zf.view(2, 5, 53, 256).permute(0, 3, 1, 2).contiguous()
    elif case == "small":      # K < 128, D = 64, ragged H*W
        z = torch.randn(3, 64, 3, 7, generator=g)
        cb = torch.randn(100, 64, generator=g)
    return z.contiguous(), cb.contiguous()


GPT_SMALL = dict(vocab_size=128, block_size=266, n_layer=2, n_head=2, n_embd=128, class_size=8, n_unmasked=0,
                 last_linear=None)
GPT_SMALL_UNMASKED = dict(vocab_size=128, block_size=266, n_layer=2, n_head=2, n_embd=128, class_size=0,
                          n_unmasked=266, last_linear=256)


def gpt_inputs(B, T, vocab, class_size, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, vocab, (B, T), generator=g)
    c = torch.randint(0, max(class_size, 1), (B, 1), generator=g)
    return x, c


def ns(cfg):
    return argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, **cfg)


def main():
    vq, gpt = ref_shim.import_reference()
    out = HERE

    # ------------------------------------------------------------------ quantiser
    for case in ("trained", "default_init", "ties", "small"):
        z, cb = vq_inputs(case)
        K, D = cb.shape
        m = vq.VectorQuantizer(K, D, 0.25)
        m._embedding.weight.data.copy_(cb)
        loss, quant, (perp, enc, idx) = m(z)
        entry = m.get_codebook_entry(idx.reshape(-1), (z.shape[0], z.shape[2], z.shape[3], D))
        entry_flat = m.get_codebook_entry(idx.reshape(-1)[:7], None)
        np.savez_compressed(os.path.join(out, "vq_%s.npz" % case),
                            idx=idx.numpy().astype(np.int16), loss=loss.numpy(), perplexity=perp.numpy(),
                            quantized=quant.numpy(), enc_rowsum=enc.sum(1).numpy(), enc_colsum=enc.sum(0).numpy(),
                            entry=entry.numpy(), entry_flat=entry_flat.numpy())
        print("vq", case, "loss", float(loss), "perplexity", float(perp))

    # ------------------------------------------------------------------ minGPT (small configs, exhaustive outputs)
    for name, cfg, B, T in (("small", GPT_SMALL, 3, 265), ("small_short", GPT_SMALL, 2, 17)):
        sd = synthetic.synthetic_gpt_state_dict(cfg, seed=101, perturb=True, with_mask=True)
        m = gpt.GPTClass(ns(cfg)).eval()
        m.load_state_dict(sd)
        x, c = gpt_inputs(B, T, cfg["vocab_size"], cfg["class_size"], seed=7)
        logits, _, att = m(x[:, :-1], c)
        np.savez_compressed(os.path.join(out, "gpt_%s.npz" % name), logits=logits.numpy(), att=att.numpy().astype(np.float32))
        print("gpt", name, logits.shape, att.shape)

    # unmasked encoder-style GPT with float prefix embedding and last_linear (GPT-VAE pieces)
    cfg = GPT_SMALL_UNMASKED
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=102, perturb=True, with_mask=True)
    m = gpt.GPT(ns(cfg), n_unmasked=cfg["n_unmasked"], last_linear=cfg["last_linear"]).eval()
    m.load_state_dict(sd)
    x, _ = gpt_inputs(2, 40, cfg["vocab_size"], 0, seed=8)
    emb = torch.randn(2, 1, cfg["n_embd"], generator=torch.Generator().manual_seed(9)) * 0.1
    logits, _, att = m(x, embeddings=emb)
    np.savez_compressed(os.path.join(out, "gpt_small_unmasked.npz"), logits=logits.numpy(), att=att.numpy())
    print("gpt unmasked", logits.shape)

    # greedy + step logits on the small config through Lit_minGPT.sample's loop (restated by hand on
    # the reference GPTClass because constructing Lit_minGPT needs the dataset on disk)
    cfg = GPT_SMALL
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=101, perturb=True, with_mask=True)
    # sharpen the output distribution so greedy decoding is decisive
    sd["head.weight"] = sd["head.weight"] * 8.0
    m = gpt.GPTClass(ns(cfg)).eval()
    m.load_state_dict(sd)
    lit = gpt.Lit_minGPT.__new__(gpt.Lit_minGPT)      # reference class, bypassing the dataset-loading __init__
    torch.nn.Module.__init__(lit)
    lit.transformer = m
    lit.pkeep = 1.0
    c = torch.tensor([[3], [5]])
    x0 = torch.zeros(2, 0, dtype=torch.long)
    xs, att = gpt.Lit_minGPT.sample(lit, x0, c, steps=265, temperature=1.0, sample=False, top_k=None)
    logits_tf, _ = gpt.Lit_minGPT.forward(lit, xs, c)
    np.savez_compressed(os.path.join(out, "gpt_small_greedy.npz"), tokens=xs.numpy().astype(np.int16),
                        att=att.numpy(), logits_tf=logits_tf.numpy())
    print("greedy tokens", xs[0, :12].tolist())
    # half-prompt continuation (log_images: x[:, :132] + 133 steps), top_k path
    xs2, att2 = gpt.Lit_minGPT.sample(lit, xs[:, :132], c, steps=133, temperature=0.7, sample=False, top_k=100)
    np.savez_compressed(os.path.join(out, "gpt_small_greedy_half.npz"), tokens=xs2.numpy().astype(np.int16), att_last_rows=att2[:, :, -3:].numpy())
    # top_k_logits
    lg = torch.randn(5, 128, generator=torch.Generator().manual_seed(5))
    lg[0, 10] = lg[0, 20]  # a tie
    np.savez_compressed(os.path.join(out, "topk.npz"), logits=lg.numpy(), out100=gpt.Lit_minGPT.top_k_logits(lit, lg, 100).numpy(),
                        out1=gpt.Lit_minGPT.top_k_logits(lit, lg, 1).numpy())
    # code_reader / make_idx
    fwd, bwd = gpt.Lit_minGPT.make_idx(lit, 5, 53)
    np.savez_compressed(os.path.join(out, "code_reader.npz"), fwd=fwd.numpy(), bwd=bwd.numpy())

    # ------------------------------------------------------------------ minGPT full VAS config
    cfg = synthetic.GPT_VAS
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=True, with_mask=True)
    m = gpt.GPTClass(ns(cfg)).eval()
    m.load_state_dict(sd)
    x, c = gpt_inputs(2, 265, 128, 8, seed=0)
    logits, _, att = m(x[:, :-1], c)
    # yardstick: the reference's own error when it computes in bf16 (torch.autocast) instead of fp32
    with torch.autocast("cpu", dtype=torch.bfloat16):
        logits_bf, _, att_bf = m(x[:, :-1], c)
    eb = (logits_bf.float() - logits).abs()
    ea = (att_bf.float() - att).abs()
    np.savez_compressed(os.path.join(out, "gpt_vas.npz"), logits=logits.numpy(), att_rows=att[:, :, ::33].numpy(),
                        autocast_logit_err_max=eb.max().numpy(), autocast_logit_err_rms=eb.pow(2).mean().sqrt().numpy(),
                        autocast_att_err_max=ea.max().numpy())
    print("autocast yardstick: logits max %.4f rms %.4f att max %.5f" % (eb.max(), eb.pow(2).mean().sqrt(), ea.max()))
    print("gpt vas", logits.shape, float(logits.std()))
    del m, sd

    # ------------------------------------------------------------------ VQVAE encoder / decoder
    sd = synthetic.synthetic_vqvae_state_dict(128, 256, seed=783435, perturb=True)
    m = vq.LitVQVAE(128, 256).eval()
    missing = m.load_state_dict(sd, strict=False)
    assert all(k.startswith("discriminator.") for k in missing.missing_keys), missing.missing_keys
    assert not missing.unexpected_keys, missing.unexpected_keys
    g = torch.Generator().manual_seed(21)
    codes = torch.randint(0, 128, (1, 265), generator=g)
    quant = m._vq_vae.get_codebook_entry(codes.reshape(-1), (1, 5, 53, 256))
    mel = m.decode(quant)
    melin = torch.rand(1, 1, 80, 848, generator=g) * 2 - 1
    z = m.encode(melin)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        mel_bf = m.decode(quant)
        z_bf = m.encode(melin)
    em = (mel_bf.float() - mel).abs()
    ez = (z_bf.float() - z).abs()
    print("autocast yardstick: mel max %.4f rms %.4f ; z max %.4f rms %.4f" % (em.max(), em.pow(2).mean().sqrt(), ez.max(), ez.pow(2).mean().sqrt()))
    np.savez_compressed(os.path.join(out, "vqvae.npz"), mel=mel.numpy(), z=z.numpy(),
                        autocast_mel_err_max=em.max().numpy(), autocast_mel_err_rms=em.pow(2).mean().sqrt().numpy(),
                        autocast_z_err_max=ez.max().numpy(), autocast_z_err_rms=ez.pow(2).mean().sqrt().numpy())
    print("vqvae mel", mel.shape, float(mel.std()), "z", z.shape, float(z.std()))


if __name__ == "__main__":
    main()
