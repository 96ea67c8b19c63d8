"""Shared helpers for the parity tests."""
import argparse
import os

import numpy as np
import torch

from melspec_gpt_vqvae_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def gpt_args(cfg):
    return argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, **cfg)


def make_gpt(cfg, sd, device="cuda"):
    """Drop-in GPTClass / GPT loaded with `sd` (reference state_dict layout), eval mode, on the GPU."""
    from melspec_gpt_vqvae_b200.transformer.minGPT import GPT, GPTClass
    if cfg.get("class_size", 0):
        m = GPTClass(gpt_args(cfg))
    else:
        m = GPT(gpt_args(cfg), n_unmasked=cfg.get("n_unmasked", 0), last_linear=cfg.get("last_linear"))
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    assert all(k.endswith("attn.mask") for k in missing.missing_keys), missing.missing_keys
    return m.eval().to(device)


def make_vqvae(sd, device="cuda"):
    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE
    m = LitVQVAE(128, 256)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert all(k.startswith("discriminator.") for k in missing.missing_keys)
    return m.eval().to(device)


def err_stats(a, b):
    d = (a.double() - b.double()).abs()
    return float(d.max()), float(d.pow(2).mean().sqrt())
