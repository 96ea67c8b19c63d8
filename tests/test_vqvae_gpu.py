"""GPU parity: VQVAE decoder / encoder (drop-in LitVQVAE -> C ABI) against the reference's golden
outputs and the fp32 oracle.  bf16 compute: tolerance = 2x the unmodified reference's own error
under torch.autocast(bfloat16) (autocast_* in tests/golden/vqvae.npz)."""
import os

import numpy as np
import pytest
import torch

from helpers import err_stats, golden, make_vqvae
from melspec_gpt_vqvae_b200 import synthetic
from oracle import vq_oracle, vqvae_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_and_sd():
    sd = synthetic.synthetic_vqvae_state_dict(128, 256, seed=783435, perturb=True)
    return make_vqvae(sd), sd


def test_decode_codes_vs_golden(model_and_sd):
    m, sd = model_and_sd
    g = golden("vqvae.npz")
    gen = torch.Generator().manual_seed(21)
    codes = torch.randint(0, 128, (1, 265), generator=gen)
    mel = m.decode_codes(codes.cuda())
    assert mel.shape == (1, 1, 80, 848) and mel.dtype == torch.float32
    emax, erms = err_stats(mel.cpu(), torch.from_numpy(g["mel"]))
    ymax, yrms = float(g["autocast_mel_err_max"]), float(g["autocast_mel_err_rms"])
    print("decode mel err max %.4f rms %.4f (reference under bf16 autocast: max %.4f rms %.4f; mel std %.3f)" %
          (emax, erms, ymax, yrms, float(g["mel"].std())))
    assert erms <= 2 * yrms and emax <= 2 * ymax
    # decode(quant) path (LitVQVAE.decode signature) gives the same mel as the fused-gather path
    quant = m._vq_vae.get_codebook_entry(codes.reshape(-1).cuda(), (1, 5, 53, 256))
    mel2 = m.decode(quant)
    e2, r2 = err_stats(mel2.cpu(), torch.from_numpy(g["mel"]))
    assert r2 <= 2 * yrms and e2 <= 2 * ymax


def test_decode_graph_replay_tracks_batch_size_and_weights():
    """The code path of the decoder replays a CUDA graph from its third call at a batch size on (first call eager, second
    captures).  The replayed result must equal the eager one, survive a change of batch size in between (new capture) and
    follow in-place weight updates (the graph holds pointers to the packed weights, which are rewritten in place)."""
    sd = synthetic.synthetic_vqvae_state_dict(128, 256, seed=11, perturb=True)
    m = make_vqvae(sd)
    gen = torch.Generator().manual_seed(9)
    codes = torch.randint(0, 128, (4, 265), generator=gen).cuda()
    eager = m.decode_codes(codes).cpu()                      # call 1 at B=4: eager
    captured = m.decode_codes(codes).cpu()                   # call 2: capture + launch
    replayed = m.decode_codes(codes).cpu()                   # call 3: replay
    assert torch.equal(eager, captured) and torch.equal(eager, replayed)
    other = torch.randint(0, 128, (4, 265), generator=gen).cuda()
    assert not torch.equal(m.decode_codes(other).cpu(), eager), "replay must read the new indices"
    two = m.decode_codes(codes[:2]).cpu()                    # another batch size drops the graph
    assert torch.equal(two, eager[:2])
    for _ in range(3):                                       # eager, capture, replay at B=4 again
        assert torch.equal(m.decode_codes(codes).cpu(), eager)
    with torch.no_grad():
        m._decoder.conv_out.bias.add_(0.5)                   # the tail adds this bias to every mel bin
    shifted = m.decode_codes(codes).cpu()
    assert torch.allclose(shifted, eager + 0.5, atol=1e-5), "weight update not visible through the replayed graph"
    with pytest.raises(RuntimeError):                        # the index check still fires on the replay path
        m.decode_codes(torch.full((4, 265), 128, dtype=torch.long, device="cuda"))
    assert torch.equal(m.decode_codes(codes).cpu(), shifted)


def test_decode_batch_independence_and_oracle(model_and_sd):
    m, sd = model_and_sd
    gen = torch.Generator().manual_seed(5)
    codes = torch.randint(0, 128, (3, 265), generator=gen)
    mel = m.decode_codes(codes.cuda()).cpu()
    ref = vqvae_oracle.decode_codes(sd, codes, 3)
    emax, erms = err_stats(mel, ref)
    print("decode B=3 vs oracle: max %.4f rms %.4f" % (emax, erms))
    assert erms <= 0.035 and emax <= 0.2
    one = m.decode_codes(codes[1:2].cuda()).cpu()
    # a clip's mel does not depend on its batch neighbours: every reduction (GroupNorm statistics included) is
    # per image and deterministic, so the result is bit-identical
    dmax, drms = err_stats(one, mel[1:2])
    print("batch independence: max %.6f rms %.6f" % (dmax, drms))
    assert dmax == 0.0
    again = m.decode_codes(codes.cuda()).cpu()
    assert torch.equal(again, mel), "decode is not deterministic run to run"
    with pytest.raises(RuntimeError):
        m.decode_codes(torch.full((1, 265), 128, dtype=torch.long, device="cuda"))


def test_encode_vs_golden(model_and_sd):
    m, sd = model_and_sd
    g = golden("vqvae.npz")
    gen = torch.Generator().manual_seed(21)
    _ = torch.randint(0, 128, (1, 265), generator=gen)
    melin = torch.rand(1, 1, 80, 848, generator=gen) * 2 - 1
    z = m.encode(melin.cuda())
    assert z.shape == (1, 256, 5, 53)
    emax, erms = err_stats(z.cpu(), torch.from_numpy(g["z"]))
    ymax, yrms = float(g["autocast_z_err_max"]), float(g["autocast_z_err_rms"])
    print("encode z err max %.4f rms %.4f (reference under bf16 autocast: max %.4f rms %.4f)" % (emax, erms, ymax, yrms))
    assert erms <= 2 * yrms and emax <= 2 * ymax


def test_extract_codes_end_to_end(tmp_path, model_and_sd):
    """config 2 in miniature: synthetic *_mel.npy files -> codes_10s/*_code.npy (int64 5x53), skip-if-exists,
    damaged-file isolation; index agreement with the fp32 oracle encoder + exact quantiser is REPORTED
    (bf16 encoder moves z by ~1e-2, which flips near-ties: SURVEY section 7 'hard parts')."""
    from melspec_gpt_vqvae_b200.feature_extraction import extract_codes as ec
    m, sd = model_and_sd
    sd_t = dict(sd)
    gcb = torch.Generator().manual_seed(77)
    sd_t["_vq_vae._embedding.weight"] = torch.randn(128, 256, generator=gcb) * 0.2     # trained-scale codebook
    m2 = make_vqvae(sd_t)
    d = tmp_path / "features" / "dog" / "melspec_10s_22050hz"
    d.mkdir(parents=True)
    rs = np.random.RandomState(0)
    paths = []
    for i in range(5):
        p = str(d / ("clip%02d_mel.npy" % i))
        np.save(p, rs.rand(80, 860).astype(np.float32))
        paths.append(p)
    bad = str(d / "broken_mel.npy")
    np.save(bad, rs.rand(80, 100).astype(np.float32))     # too short to crop -> "is damaged"
    tr = ec.Crop([80, 848], False)
    n = ec.get_codes_batch(paths + [bad], torch.device("cuda"), 848, m2, tr, batch_size=4)
    assert n == 5
    outs = [str(tmp_path / "features" / "dog" / "codes_10s" / ("clip%02d_mel_code.npy" % i)) for i in range(5)]
    agree = []
    for p, o in zip(paths, outs):
        codes = np.load(o)
        assert codes.dtype == np.int64 and codes.shape == (5, 53)
        mel = 2 * tr(np.load(p).astype(np.float32)) - 1
        z_ref = vqvae_oracle.encode(sd_t, torch.from_numpy(mel)[None, None])
        idx_ref, _ = vq_oracle.argmin_exact(z_ref.numpy(), sd_t["_vq_vae._embedding.weight"].numpy())
        agree.append(float((codes.reshape(-1) == idx_ref).mean()))
    print("extract_codes index agreement with fp32-encoder oracle per clip:", ["%.3f" % a for a in agree])
    assert min(agree) > 0.80
    assert not os.path.exists(str(tmp_path / "features" / "dog" / "codes_10s" / "broken_mel_code.npy"))
    # single-file entry point with the reference signature; second call is a no-op ("file exists")
    os.remove(outs[0])
    ec.get_codes(paths[0], torch.device("cuda"), 848, m2, tr)
    first = np.load(outs[0])
    mtime = os.path.getmtime(outs[0])
    ec.get_codes(paths[0], torch.device("cuda"), 848, m2, tr)
    assert os.path.getmtime(outs[0]) == mtime and first.shape == (5, 53)
    # centre crop = columns 6..853 (albumentations.CenterCrop semantics)
    a = np.arange(80 * 860, dtype=np.float32).reshape(80, 860)
    assert np.array_equal(tr(a), a[:, 6:854])
