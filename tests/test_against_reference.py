"""Optional cross-checks of the oracle against the LIVE, unmodified reference (imported from /root/reference through
tests/golden/ref_shim.py) on inputs that are NOT in the committed fixtures.  Skipped where the reference is absent
(the GPU box); the committed golden fixtures (tests/golden/*.npz) remain the portable pin."""
import argparse

import numpy as np
import pytest
import torch

import ref_shim
from melspec_gpt_vqvae_b200 import synthetic
from oracle import gpt_oracle, gpt_vae_oracle, vq_oracle

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not present")


@pytest.mark.parametrize("seed,K,shape", [(1, 128, (2, 256, 3, 7)), (2, 37, (1, 64, 5, 53)), (3, 1024, (3, 256, 2, 9))])
def test_vq_oracle_vs_live_reference(seed, K, shape):
    vq, _ = ref_shim.import_reference()
    g = torch.Generator().manual_seed(seed)
    D = shape[1]
    z = torch.randn(*shape, generator=g) * 0.3
    cb = torch.randn(K, D, generator=g) * 0.3
    m = vq.VectorQuantizer(K, D, 0.25)
    m._embedding.weight.data.copy_(cb)
    with torch.no_grad():
        loss, quant, (perp, enc, idx) = m(z)
    o_idx, _ = vq_oracle.argmin_exact(z.numpy(), cb.numpy())
    ref_idx = idx.numpy().reshape(-1)
    dist = vq_oracle.distances_exact(z.numpy(), cb.numpy())
    mism = vq_oracle.classify_mismatches(dist, o_idx, ref_idx)
    # the reference computes distances with a BLAS matmul (different summation order): only near-ties may differ
    assert mism["n_real"] == 0, mism
    o_loss, o_quant, (o_perp, o_enc, o_i) = vq_oracle.forward_numpy(z.numpy(), cb.numpy(), 0.25, indices=ref_idx)
    np.testing.assert_allclose(float(loss), float(o_loss), rtol=1e-5)
    np.testing.assert_allclose(float(perp), float(o_perp), rtol=1e-4)
    np.testing.assert_array_equal(quant.numpy(), o_quant)
    np.testing.assert_array_equal(enc.numpy(), o_enc)


def test_gpt_oracle_vs_live_reference_random_config():
    _, gpt = ref_shim.import_reference()
    cfg = dict(vocab_size=50, block_size=20, n_layer=1, n_head=2, n_embd=128, class_size=3, n_unmasked=0, last_linear=None)
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=11, perturb=True, with_mask=True)
    m = gpt.GPTClass(argparse.Namespace(embd_pdrop=0.1, resid_pdrop=0.1, attn_pdrop=0.1, **cfg)).eval()
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(12)
    x = torch.randint(0, 50, (4, 19), generator=g)
    c = torch.randint(0, 3, (4, 1), generator=g)
    with torch.no_grad():
        logits, _, att = m(x, c)
    o_logits, _, o_att = gpt_oracle.gptclass_forward(sd, gpt_oracle.GPTCfg(**cfg), x, c)
    np.testing.assert_allclose(o_logits.numpy(), logits.numpy(), atol=2e-5)
    np.testing.assert_allclose(o_att.numpy(), att.numpy(), atol=1e-6)
    # top-k rule with ties
    lg = torch.randn(3, 50, generator=g)
    lg[1, 4] = lg[1, 9]
    lit = gpt.Lit_minGPT.__new__(gpt.Lit_minGPT)
    for k in (1, 7, 50):
        assert torch.equal(gpt.Lit_minGPT.top_k_logits(lit, lg, k), gpt_oracle.top_k_logits(lg, k))


def test_gpt_vae_oracle_vs_live_reference_random_config():
    ref_shim.import_reference()
    import importlib
    import os
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    try:
        enc_mod = importlib.import_module("transformer.encoders")
        dec_mod = importlib.import_module("transformer.decoders")
    finally:
        os.chdir(cwd)
    cfg = dict(vocab_size=40, block_size=12, n_layer=1, n_head=2, n_embd=128)
    args = argparse.Namespace(embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, fix_var=-1.0, device="cpu", **cfg)
    enc = enc_mod.GPTEncoder(args, n_unmasked=12, last_linear=256).eval()
    dec = dec_mod.GPTDecoder(args, block_size=13).eval()
    esd = synthetic.synthetic_gpt_state_dict(dict(cfg, class_size=0, n_unmasked=12, last_linear=256), seed=21, with_mask=True)
    dsd = synthetic.synthetic_gpt_state_dict(dict(cfg, class_size=0, n_unmasked=0, last_linear=None, block_size=13), seed=22,
                                             with_mask=True)
    enc.transformer.load_state_dict(esd)
    dec.transformer.load_state_dict(dsd)
    g = torch.Generator().manual_seed(23)
    x = torch.randint(0, 40, (5, 12), generator=g)
    z = torch.randn(5, 1, 128, generator=g)
    with torch.no_grad():
        mean, logvar, _ = enc(x)
        rec = dec.reconstruct_error(x, z)
    ecfg = gpt_vae_oracle.encoder_cfg(**cfg)
    dcfg = gpt_vae_oracle.decoder_cfg(**cfg)
    o_mean, o_logvar, _ = gpt_vae_oracle.encoder_forward(esd, ecfg, x)
    np.testing.assert_allclose(o_mean.numpy(), mean.numpy(), atol=2e-5)
    np.testing.assert_allclose(o_logvar.numpy(), logvar.numpy(), atol=2e-5)
    np.testing.assert_allclose(gpt_vae_oracle.reconstruct_error(dsd, dcfg, x, z).numpy(), rec.numpy(), rtol=1e-5)
