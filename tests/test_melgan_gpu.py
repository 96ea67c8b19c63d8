"""GPU parity of the MelGAN vocoder drop-in: vocoder/modules.py:Generator through the C ABI
against the golden waveform of the unmodified reference class and against the fp32 oracle.  fp32 compute: the tolerance is
2e-4 absolute on a waveform in [-1, 1] (summation order only)."""
import numpy as np
import pytest
import torch

from helpers import golden
from make_golden_melgan import SMALL, melgan_inputs
from melspec_gpt_vqvae_b200 import synthetic
from oracle import melgan_oracle

pytestmark = pytest.mark.gpu
TOL = 2e-4


def _gen(cfg, seed):
    from melspec_gpt_vqvae_b200.vocoder.modules import Generator
    g = Generator(cfg["n_mel"], cfg["ngf"], cfg["n_residual_layers"])
    sd = synthetic.synthetic_melgan_state_dict(seed=seed, **cfg)
    g.load_state_dict(sd, strict=True)
    return g.eval().cuda(), sd


def test_melgan_matches_reference_golden():
    g = golden("melgan_small.npz")
    gen, sd = _gen(SMALL, 410)
    mel = melgan_inputs()
    wave = gen(mel.cuda())
    assert wave.shape == (2, 1, 37 * 256) and gen.hop_length == int(g["hop_length"]) == 256
    err = float((wave.cpu() - torch.from_numpy(g["wave"])).abs().max())
    print("melgan vs reference golden: max abs err %.2e (waveform std %.3f; reference's own bf16-autocast error %.1e)"
          % (err, float(g["wave"].std()), float(g["bf16_max"])))
    assert err < TOL
    assert float(wave.abs().max()) <= 1.0
    # same call again (cached handle), and after an in-place weight change (the packed copy must follow)
    assert torch.equal(gen(mel.cuda()), wave)
    with torch.no_grad():
        gen.model["1"].bias.add_(0.25)
    sd2 = dict(sd)
    sd2["model.1.bias"] = sd["model.1.bias"] + 0.25
    ref2 = melgan_oracle.generator_forward(sd2, mel)
    assert float((gen(mel.cuda()).cpu() - ref2).abs().max()) < TOL


@pytest.mark.parametrize("cfg,B,T", [(dict(n_mel=80, ngf=32, n_residual_layers=3), 1, 53),     # the shipped vggsound vocoder
                                     (dict(n_mel=80, ngf=8, n_residual_layers=2), 3, 4),       # shortest legal clip
                                     (dict(n_mel=40, ngf=16, n_residual_layers=1), 2, 129)])   # other widths, T past one tile
def test_melgan_configs_against_oracle(cfg, B, T):
    gen, sd = _gen(cfg, 500 + T)
    gtor = torch.Generator().manual_seed(T)
    mel = torch.rand(B, cfg["n_mel"], T, generator=gtor)
    ref = melgan_oracle.generator_forward(sd, mel, cfg["n_residual_layers"])
    wave = gen(mel.cuda())
    assert wave.shape == ref.shape == (B, 1, 256 * T)
    err = float((wave.cpu() - ref).abs().max())
    print("melgan %s B=%d T=%d: max abs err %.2e, std %.3f" % (cfg, B, T, err, float(ref.std())))
    assert err < TOL
    # clips are independent: a batch equals its clips run one by one
    one = gen(mel[1:2].cuda()) if B > 1 else wave
    assert torch.equal(one, wave[1:2]) if B > 1 else True


def test_melgan_full_clip_properties_and_errors():
    """A full 848-frame clip (217 088 samples): finite, bounded by tanh, time-local (a change in the last mel frames
    cannot reach the first samples: receptive field), and the error behaviour of the boundary."""
    cfg = dict(n_mel=80, ngf=32, n_residual_layers=3)
    gen, sd = _gen(cfg, 7)
    gtor = torch.Generator().manual_seed(1)
    mel = torch.rand(2, 80, 848, generator=gtor).cuda()
    w = gen(mel)
    assert w.shape == (2, 1, 217088) and bool(torch.isfinite(w).all()) and float(w.abs().max()) <= 1.0
    mel2 = mel.clone()
    mel2[:, :, 800:] = 0.0
    w2 = gen(mel2)
    assert torch.equal(w2[..., :150000], w[..., :150000]) and not torch.equal(w2[..., 210000:], w[..., 210000:])
    with pytest.raises(RuntimeError):
        gen(mel.cpu())
    with pytest.raises(RuntimeError):
        gen(mel[:, :40])
    with pytest.raises(RuntimeError):
        gen(mel[:, :, :3].contiguous())
    assert gen(mel[:0]).shape == (0, 1, 217088)
