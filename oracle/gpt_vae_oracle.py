"""ORACLE -- test infrastructure, not product code.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

CPU fp32 restatement of the GPT-VAE pieces of karchkha/MelSpec_GPT_VQVAE on top of gpt_oracle.gpt_forward:
  GPTEncoder.forward / encode (KL) / eval_inference_dist   transformer/encoders.py:21-45, 65-83, 110-138
  GPTDecoder.forward / reconstruct_error / sample          transformer/decoders.py:23-39, 41-70, 91-123
  GPT_VAE.loss                                             transformer/Lit_GPT_VAE.py:179-195
Functional form on plain state_dicts (encoder GPT: n_unmasked = block_size, head -> 2*n_embd; decoder GPT:
block_size + 1 positions, z as the float prefix embedding).  Parity pin: tests/golden/gpt_vae_small.npz holds the
outputs of the UNMODIFIED reference classes (tests/golden/make_golden_vae.py); tests/test_oracle_cpu.py checks this
file against it.
"""
import math

import torch
import torch.nn.functional as F

from . import gpt_oracle


def encoder_cfg(vocab_size, block_size, n_layer, n_head, n_embd):
    return gpt_oracle.GPTCfg(vocab_size=vocab_size, block_size=block_size, n_layer=n_layer, n_head=n_head, n_embd=n_embd,
                             class_size=0, n_unmasked=block_size, last_linear=2 * n_embd)


def decoder_cfg(vocab_size, block_size, n_layer, n_head, n_embd):
    return gpt_oracle.GPTCfg(vocab_size=vocab_size, block_size=block_size + 1, n_layer=n_layer, n_head=n_head,
                             n_embd=n_embd, class_size=0, n_unmasked=0, last_linear=None)


@torch.no_grad()
def encoder_forward(sd, cfg, x, fix_var=-1.0):
    """GPTEncoder.forward (encoders.py:21-45): -> mean, logvar, att."""
    logits, _, att = gpt_oracle.gpt_forward(sd, cfg, x)                                    # :33
    mean, logvar = logits[:, -1, :].chunk(2, -1)                                           # :35-37
    if fix_var > 0:
        logvar = torch.full_like(mean, math.log(fix_var))                                  # :40-41
    return mean, logvar, att


def kl_to_standard_normal(mean, logvar):
    """encoders.py:81"""
    return 0.5 * (mean.pow(2) + logvar.exp() - logvar - 1).sum(dim=1)


def eval_inference_dist(mean, logvar, z):
    """log q(z|x), z (B, ns, nz) -> (B, ns)   (encoders.py:121-138)"""
    nz = z.size(2)
    mu, lv = mean.unsqueeze(1), logvar.unsqueeze(1)
    dev = z - mu
    return -0.5 * ((dev ** 2) / lv.exp()).sum(dim=-1) - 0.5 * (nz * math.log(2 * math.pi) + lv.sum(-1))


@torch.no_grad()
def decoder_forward(sd, cfg, x, z):
    """GPTDecoder.forward (decoders.py:23-39): logits row i = p(x_i | x_<i, z)."""
    logits, _, _ = gpt_oracle.gpt_forward(sd, cfg, x[:, :-1], embeddings=z)               # :33
    cond_size = z.size(-2)                                                                 # :35
    return logits[:, cond_size - 1:], x


@torch.no_grad()
def reconstruct_error(sd, cfg, x, z):
    """GPTDecoder.reconstruct_error (decoders.py:41-70), n_sample == 1 path: (B, 1)."""
    B, T = x.size()
    assert z.size(1) == 1
    logits, tgt = decoder_forward(sd, cfg, x, z)
    loss = F.cross_entropy(logits.reshape(-1, logits.size(2)), tgt.reshape(-1), reduction="none")   # :64-65
    return loss.view(B, 1, -1).sum(-1)                                                     # :68


@torch.no_grad()
def decoder_sample(sd, cfg, x, z, steps, temperature=1.0, sample=False, top_k=None, generator=None):
    """GPTDecoder.sample (decoders.py:91-123): full forward per step, z as the prefix."""
    att = None
    for _ in range(steps):
        assert x.size(1) + z.size(-2) <= cfg.block_size                                    # :102
        logits, _, att = gpt_oracle.gpt_forward(sd, cfg, x, embeddings=z)                  # :107
        logits = logits[:, -1, :] / temperature                                            # :110
        if top_k is not None:
            logits = gpt_oracle.top_k_logits(logits, top_k)                                # :113
        probs = F.softmax(logits, dim=-1)                                                  # :115
        if sample:
            ix = torch.multinomial(probs, num_samples=1, generator=generator)              # :118
        else:
            _, ix = torch.topk(probs, k=1, dim=-1)                                         # :120
        x = torch.cat((x, ix), dim=1)                                                      # :122
    return x, att


@torch.no_grad()
def vae_loss(enc_sd, enc_cfg, dec_sd, dec_cfg, x, z, kl_weight):
    """GPT_VAE.loss (Lit_GPT_VAE.py:179-195) for a given posterior sample z (B, 1, nz)."""
    mean, logvar, _ = encoder_forward(enc_sd, enc_cfg, x)
    kl = kl_to_standard_normal(mean, logvar)
    rec = reconstruct_error(dec_sd, dec_cfg, x, z).mean(dim=1)
    return rec + kl_weight * kl, rec, kl
