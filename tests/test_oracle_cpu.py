"""CPU: the oracle (oracle/*.py, oracle/vq_oracle.c) against the golden fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from make_golden import GPT_SMALL, GPT_SMALL_UNMASKED, gpt_inputs, vq_inputs
from melspec_gpt_vqvae_b200 import synthetic
from make_golden_vae import GPT_VAE_SMALL, vae_inputs, vae_state_dicts
from oracle import gpt_oracle, gpt_vae_oracle, vq_oracle, vqvae_oracle


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("case", ["trained", "default_init", "ties", "small"])
def test_vq_oracle_vs_reference(golden_dir, case):
    g = load(golden_dir, "vq_%s.npz" % case)
    z, cb = vq_inputs(case)
    z, cb = z.numpy(), cb.numpy()
    ref_idx = g["idx"].astype(np.int64).reshape(-1)
    # (1) fixed-order C restatement: equal to the reference except classified near-ties
    idx, dmin = vq_oracle.argmin_exact(z, cb)
    dist = vq_oracle.distances_exact(z, cb)
    assert np.array_equal(dist.argmin(1), idx)                      # first-index argmin
    assert np.array_equal(dist[np.arange(idx.size), idx], dmin)
    rep = vq_oracle.classify_mismatches(dist, idx, ref_idx, ulps=16)
    print(case, rep)
    assert rep["n_real"] == 0, rep
    if case == "trained":
        assert rep["n_mismatch"] == 0, rep
    if case == "ties":
        # exact duplicates (code j and j+64): the lower index must win
        assert (idx < 64).all()
    # (2) formula restatement in numpy: remaining outputs on the reference's own indices
    loss, quant, (perp, enc, enc_idx) = vq_oracle.forward_numpy(z, cb, 0.25, indices=ref_idx)
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-6)
    np.testing.assert_allclose(perp, g["perplexity"], rtol=1e-5)
    np.testing.assert_allclose(quant, g["quantized"], rtol=0, atol=1e-7)   # straight-through rounding <= 1 ulp
    np.testing.assert_array_equal(enc.sum(1), g["enc_rowsum"])
    np.testing.assert_array_equal(enc.sum(0), g["enc_colsum"])
    # and with its own argmin
    _, _, (_, _, own_idx) = vq_oracle.forward_numpy(z, cb, 0.25)
    rep2 = vq_oracle.classify_mismatches(dist, own_idx.reshape(-1), ref_idx, ulps=16)
    assert rep2["n_real"] == 0, rep2
    # (3) get_codebook_entry: pure gather, exact
    B, D = z.shape[0], z.shape[1]
    entry = vq_oracle.get_codebook_entry_numpy(ref_idx, cb, (B, z.shape[2], z.shape[3], D))
    np.testing.assert_array_equal(entry, g["entry"])
    np.testing.assert_array_equal(vq_oracle.get_codebook_entry_numpy(ref_idx[:7], cb, None), g["entry_flat"])


def test_vq_oracle_empty_and_single():
    cb = np.random.RandomState(0).randn(128, 256).astype(np.float32)
    idx, dmin = vq_oracle.argmin_exact(np.zeros((0, 256, 5, 53), np.float32), cb)
    assert idx.shape == (0,)
    z = cb[5].reshape(1, 256, 1, 1)
    idx, dmin = vq_oracle.argmin_exact(z, cb)
    assert idx.tolist() == [5]


@pytest.mark.parametrize("name,B,T", [("small", 3, 265), ("small_short", 2, 17)])
def test_gpt_oracle_small(golden_dir, name, B, T):
    g = load(golden_dir, "gpt_%s.npz" % name)
    cfg = gpt_oracle.GPTCfg(**GPT_SMALL)
    sd = synthetic.synthetic_gpt_state_dict(GPT_SMALL, seed=101, perturb=True)
    x, c = gpt_inputs(B, T, 128, 8, seed=7)
    logits, _, att = gpt_oracle.gptclass_forward(sd, cfg, x[:, :-1], c)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(att.numpy(), g["att"], rtol=1e-4, atol=1e-6)


def test_gpt_oracle_unmasked_prefix_embedding(golden_dir):
    g = load(golden_dir, "gpt_small_unmasked.npz")
    cfg = gpt_oracle.GPTCfg(**GPT_SMALL_UNMASKED)
    sd = synthetic.synthetic_gpt_state_dict(GPT_SMALL_UNMASKED, seed=102, perturb=True)
    x, _ = gpt_inputs(2, 40, 128, 0, seed=8)
    emb = torch.randn(2, 1, 128, generator=torch.Generator().manual_seed(9)) * 0.1
    logits, _, att = gpt_oracle.gpt_forward(sd, cfg, x, embeddings=emb)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(att.numpy(), g["att"], rtol=1e-4, atol=1e-6)


def test_gpt_oracle_greedy_sample(golden_dir):
    g = load(golden_dir, "gpt_small_greedy.npz")
    cfg = gpt_oracle.GPTCfg(**GPT_SMALL)
    sd = synthetic.synthetic_gpt_state_dict(GPT_SMALL, seed=101, perturb=True)
    sd["head.weight"] = sd["head.weight"] * 8.0
    c = torch.tensor([[3], [5]])
    ks = []
    xs, att = gpt_oracle.sample(sd, cfg, torch.zeros(2, 0, dtype=torch.long), c, steps=40, callback=ks.append)
    assert ks == list(range(40))
    np.testing.assert_array_equal(xs.numpy(), g["tokens"][:, :40])
    # teacher-forced logits on the reference's tokens
    toks = torch.from_numpy(g["tokens"].astype(np.int64))
    logits, target = gpt_oracle.lit_forward(sd, cfg, toks, c)
    np.testing.assert_allclose(logits.numpy(), g["logits_tf"], rtol=1e-4, atol=1e-4)
    assert torch.equal(target, toks)
    # half-prompt continuation with temperature and top-k
    g2 = load(golden_dir, "gpt_small_greedy_half.npz")
    xs2, att2 = gpt_oracle.sample(sd, cfg, toks[:, :132], c, steps=20, temperature=0.7, top_k=100)
    np.testing.assert_array_equal(xs2.numpy(), g2["tokens"][:, :152])


def test_block_size_assert():
    cfg = gpt_oracle.GPTCfg(**GPT_SMALL)
    sd = synthetic.synthetic_gpt_state_dict(GPT_SMALL, seed=101)
    with pytest.raises(AssertionError):
        gpt_oracle.gptclass_forward(sd, cfg, torch.zeros(1, 266, dtype=torch.long), torch.zeros(1, 1, dtype=torch.long))


def test_topk_and_code_reader(golden_dir):
    g = load(golden_dir, "topk.npz")
    lg = torch.from_numpy(g["logits"])
    np.testing.assert_array_equal(gpt_oracle.top_k_logits(lg, 100).numpy(), g["out100"])
    np.testing.assert_array_equal(gpt_oracle.top_k_logits(lg, 1).numpy(), g["out1"])
    c = load(golden_dir, "code_reader.npz")
    fwd, bwd = gpt_oracle.make_idx(5, 53)
    np.testing.assert_array_equal(fwd, c["fwd"])
    np.testing.assert_array_equal(bwd, c["bwd"])
    x = np.arange(2 * 265).reshape(2, 265)
    np.testing.assert_array_equal(gpt_oracle.code_reader(gpt_oracle.code_reader(x), reverse=True), x)
    assert fwd[:6].tolist() == [0, 53, 106, 159, 212, 1]


def test_gpt_oracle_vas_full(golden_dir):
    g = load(golden_dir, "gpt_vas.npz")
    cfg = gpt_oracle.GPTCfg(**synthetic.GPT_VAS)
    sd = synthetic.synthetic_gpt_state_dict(synthetic.GPT_VAS, seed=783435, perturb=True)
    x, c = gpt_inputs(2, 265, 128, 8, seed=0)
    logits, _, att = gpt_oracle.gptclass_forward(sd, cfg, x[:, :-1], c)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(att[:, :, ::33].numpy(), g["att_rows"], rtol=1e-4, atol=1e-6)


def test_vqvae_oracle(golden_dir):
    g = load(golden_dir, "vqvae.npz")
    sd = synthetic.synthetic_vqvae_state_dict(128, 256, seed=783435, perturb=True)
    gen = torch.Generator().manual_seed(21)
    codes = torch.randint(0, 128, (1, 265), generator=gen)
    mel = vqvae_oracle.decode_codes(sd, codes, 1)
    np.testing.assert_allclose(mel.numpy(), g["mel"], rtol=1e-4, atol=1e-5)
    melin = torch.rand(1, 1, 80, 848, generator=gen) * 2 - 1
    z = vqvae_oracle.encode(sd, melin)
    np.testing.assert_allclose(z.numpy(), g["z"], rtol=1e-4, atol=1e-5)


def test_gpt_vae_oracle_vs_reference_golden(golden_dir):
    """GPTEncoder / GPTDecoder restatement against the unmodified reference classes' outputs."""
    g = load(golden_dir, "gpt_vae_small.npz")
    ecfg = gpt_vae_oracle.encoder_cfg(**GPT_VAE_SMALL)
    dcfg = gpt_vae_oracle.decoder_cfg(**GPT_VAE_SMALL)
    esd, dsd = vae_state_dicts()
    x, z, z2 = vae_inputs()
    mean, logvar, att = gpt_vae_oracle.encoder_forward(esd, ecfg, x)
    np.testing.assert_allclose(mean.numpy(), g["mean"], atol=2e-5)
    np.testing.assert_allclose(logvar.numpy(), g["logvar"], atol=2e-5)
    np.testing.assert_allclose(att[:, :, ::66].numpy(), g["att_enc_rows"], atol=1e-6)
    # fully unmasked attention: the first row attends to every position
    assert float(att[:, :, 0, -1].min()) > 0
    np.testing.assert_allclose(gpt_vae_oracle.kl_to_standard_normal(mean, logvar).numpy(), g["kl"], rtol=1e-5)
    np.testing.assert_allclose(gpt_vae_oracle.eval_inference_dist(mean, logvar, z2).numpy(), g["logq"], rtol=1e-5)
    logits, _ = gpt_vae_oracle.decoder_forward(dsd, dcfg, x, z)
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=2e-5)
    np.testing.assert_allclose(gpt_vae_oracle.reconstruct_error(dsd, dcfg, x, z).numpy(), g["rec"], rtol=1e-5)
    total, rec, kl = gpt_vae_oracle.vae_loss(esd, ecfg, dsd, dcfg, x, z, 0.5)
    np.testing.assert_allclose(total.numpy(), g["rec"][:, 0] + 0.5 * g["kl"], rtol=1e-5)
    # greedy generation from z (sharpened head), first 24 steps
    sharp = dict(dsd)
    sharp["head.weight"] = dsd["head.weight"] * 8.0
    toks, _ = gpt_vae_oracle.decoder_sample(sharp, dcfg, torch.zeros(3, 0, dtype=torch.long), z, steps=24)
    assert np.array_equal(toks.numpy(), g["tokens"][:, :24].astype(np.int64))


def test_melgan_oracle_vs_reference_golden(golden_dir):
    """MelGAN Generator restatement against the unmodified reference class (fixture: make_golden_melgan.py)."""
    from make_golden_melgan import SMALL, melgan_inputs
    from oracle import melgan_oracle
    g = load(golden_dir, "melgan_small.npz")
    sd = synthetic.synthetic_melgan_state_dict(seed=410, **SMALL)
    wave = melgan_oracle.generator_forward(sd, melgan_inputs())
    assert wave.shape == (2, 1, 37 * 256)
    np.testing.assert_allclose(wave.numpy(), g["wave"], atol=2e-6)
    assert float(wave.std()) > 0.05          # the fixture exercises the nonlinearity, not a flat line


def test_upsample_phase_decomposition_is_exact():
    """Upsample.forward (reference vqvae/big_model_attn_gan.py:182-186) = nearest 2x, then conv3x3 pad 1.  libmgv computes it
    as four 2x2 convolutions over the LOW-res tensor (upsample_phase_weights_kernel / conv_up): output pixel (2y+py, 2x+px)
    sees input rows {y-1, y} for py = 0 and {y, y+1} for py = 1 (same in x), with the filter rows that land on the same
    input row summed.  This is the index rule of the kernel restated in torch fp64 against the literal form, including
    the borders (zero padding of the upsampled image = out-of-range low-res pixels)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 5, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 6, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(4, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1)
    out = torch.zeros_like(ref)
    fold = {0: ([0], [1, 2]), 1: ([0, 1], [2])}          # phase -> filter taps folded onto low-res tap 0 and tap 1
    for py in (0, 1):
        for px in (0, 1):
            wp = torch.zeros(4, 6, 2, 2, dtype=torch.float64)
            for ty in (0, 1):
                for tx in (0, 1):
                    for ky in fold[py][ty]:
                        for kx in fold[px][tx]:
                            wp[:, :, ty, tx] += w[:, :, ky, kx]
            # low-res tap (ty, tx) reads input (y + ty - (1 - py), x + tx - (1 - px)): pad (1 - p) before, p after
            xp = F.pad(x, (1 - px, px, 1 - py, py))
            out[:, :, py::2, px::2] = F.conv2d(xp, wp, b)
    assert torch.allclose(out, ref, atol=1e-12), (out - ref).abs().max()
