"""CPU: the C-ABI library loads and exports every symbol include/mgv.h declares (no compute calls),
host-side logic (state_dict layout, Crop, sharding, error behaviour without a device)."""
import argparse
import os
import re

import numpy as np
import pytest
import torch

from melspec_gpt_vqvae_b200 import _lib, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mgv.h")).read()
    declared = set(re.findall(r"\b(mgv_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libmgv.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.mgv_version() == 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    assert lib.mgv_device_check() == 3          # MGV_ERR_DEVICE
    assert "no CPU fallback" in _lib.last_error()
    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import VectorQuantizer
    vq = VectorQuantizer(128, 256, 0.25)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq(torch.zeros(1, 256, 5, 53))
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq.get_codebook_entry(torch.zeros(4, dtype=torch.long), None)


def test_gpt_state_dict_layout_matches_reference_enumeration():
    from melspec_gpt_vqvae_b200.transformer.minGPT import GPTClass
    cfg = dict(synthetic.GPT_VAS, n_layer=2)
    m = GPTClass(argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, **cfg))
    sd = m.state_dict()
    want = synthetic.gpt_param_shapes(cfg)
    keys = [k for k in sd if not k.endswith("attn.mask")]
    assert keys == list(want), "key order / names differ"
    for k, shape in want.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    assert tuple(sd["blocks.0.attn.mask"].shape) == (1, 1, 266, 266)      # persistent buffer, like the reference
    # full VAS config: 414 keys / 302 854 144 parameters
    full = synthetic.gpt_param_shapes(synthetic.GPT_VAS)
    assert len(full) + 24 == 414
    assert sum(int(np.prod(s)) for s in full.values()) == 302854144
    # init distributions (reference _init_weights): LayerNorm (1,0), Linear bias 0, pos_emb 0
    assert float(sd["ln_f.weight"].min()) == 1.0 and float(sd["blocks.0.mlp.0.bias"].abs().max()) == 0.0
    assert float(sd["pos_emb"].abs().max()) == 0.0
    assert abs(float(sd["tok_emb.weight"].std()) - 0.02) < 2e-3


def test_vqvae_state_dict_layout():
    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE
    m = LitVQVAE(128, 256)
    sd = m.state_dict()
    want = synthetic.vqvae_param_shapes(128, 256)
    hot = [k for k in sd if not k.startswith("discriminator.")]
    assert hot == list(want)
    for k, shape in want.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    n = lambda p: sum(v.numel() for k, v in sd.items() if k.startswith(p) and "num_batches" not in k and "running" not in k)
    assert n("_encoder.") == 29295872 and n("_decoder.") == 42447489          # SURVEY appendix A [probed]
    assert n("discriminator.") == 2763585 + 0 or n("discriminator.") > 0
    assert sd["discriminator.main.0.weight"].shape == (64, 1, 4, 4) and "discriminator.main.9.running_var" in sd
    # codebook init U(+-1/K)
    assert float(sd["_vq_vae._embedding.weight"].abs().max()) <= 1 / 128


def test_reference_state_dict_keys_if_available():
    import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    vq, gpt = ref_shim.import_reference()
    ref = vq.LitVQVAE(128, 256).state_dict()
    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE
    ours = LitVQVAE(128, 256).state_dict()
    assert list(ref.keys()) == list(ours.keys())
    assert all(ref[k].shape == ours[k].shape for k in ref)
    cfg = dict(synthetic.GPT_VAS, n_layer=1)
    a = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, **cfg)
    from melspec_gpt_vqvae_b200.transformer.minGPT import GPTClass
    r, o = gpt.GPTClass(a).state_dict(), GPTClass(a).state_dict()
    assert list(r.keys()) == list(o.keys()) and all(r[k].shape == o[k].shape for k in r)


def test_crop_and_shard_host_logic():
    from melspec_gpt_vqvae_b200.feature_extraction import extract_codes as ec
    a = np.arange(80 * 860, dtype=np.float32).reshape(80, 860)
    assert np.array_equal(ec.Crop([80, 848], False)(a), a[:, 6:854])
    assert ec.Crop(None)(a) is a
    with pytest.raises(ValueError):
        ec.Crop([80, 848], False)(a[:, :100])
    paths = ["f%03d" % i for i in range(10)]
    parts = [ec.shard(paths, r, 4) for r in range(4)]
    assert sum(parts, []) == paths and max(map(len, parts)) - min(map(len, parts)) <= 1
    assert ec.shard([], 0, 2) == []
    assert ec._out_path("/x/features/dog/melspec_10s_22050hz/a_mel.npy", "codes_10s") == "/x/features/dog/codes_10s/a_mel_code.npy"


def test_code_reader_host_logic():
    from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT
    lit = Lit_minGPT.__new__(Lit_minGPT)
    torch.nn.Module.__init__(lit)
    lit.forward_shuffle_idx, lit.backward_shuffle_idx = Lit_minGPT.make_idx(lit, 5, 53)
    g = np.load(os.path.join(ROOT, "tests", "golden", "code_reader.npz"))
    assert np.array_equal(lit.forward_shuffle_idx.numpy(), g["fwd"]) and np.array_equal(lit.backward_shuffle_idx.numpy(), g["bwd"])
    x = torch.arange(2 * 265).reshape(2, 265)
    assert torch.equal(lit.code_reader(lit.code_reader(x), reverse=True), x)
    # get_x: codes (B,5,53) -> (B,265) time-major == code_reader(row-major flatten)
    lit.args = argparse.Namespace(device="cpu")
    codes = torch.arange(2 * 265).reshape(2, 5, 53)
    assert torch.equal(lit.get_x({"codes": codes}), lit.code_reader(codes.reshape(2, -1)))
    lg = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "topk.npz"))["logits"])
    assert torch.equal(lit.top_k_logits(lg, 100), torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "topk.npz"))["out100"]))


def test_gpt_vae_dropin_state_dict_keys_match_reference():
    """GPT_VAE drop-in registers exactly the reference's parameters / buffers (encoder.*, decoder.*, decoder.loss.weight)."""
    import argparse
    import os
    import numpy as np
    from melspec_gpt_vqvae_b200.transformer import GPTDecoder, GPTEncoder  # noqa: F401  (the reference's import surface)
    from melspec_gpt_vqvae_b200.transformer.Lit_GPT_VAE import GPT_VAE
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gpt_vae_small.npz"))
    args = argparse.Namespace(embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, fix_var=-1.0, device="cpu", vocab_size=128,
                              block_size=265, n_layer=2, n_head=2, n_embd=128)
    m = GPT_VAE(args)
    ours = set(m.state_dict().keys())
    ref = set(str(k) for k in g["state_dict_keys"])
    masks = {k for k in ref if k.endswith("attn.mask")}      # derived from n_unmasked inside libmgv; optional buffer
    assert ours - masks == ref - masks, (sorted(ours - ref)[:5], sorted(ref - masks - ours)[:5])
    assert m.encoder.transformer.head.weight.shape == (256, 128) and m.decoder.transformer.pos_emb.shape == (1, 266, 128)


def test_melgan_dropin_state_dict_keys_and_host_errors():
    """vocoder/modules.py:Generator drop-in: the reference's state_dict keys / shapes (so best_netG.pt loads), hop length,
    weight-norm packing, and no CPU path."""
    from melspec_gpt_vqvae_b200.vocoder.modules import Generator
    gen = Generator(80, 32, 3)
    shapes = synthetic.melgan_param_shapes(80, 32, 3)
    sd = gen.state_dict()
    assert set(sd) == set(shapes) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    assert gen.hop_length == 256
    # fresh initialisation: g = |v|, i.e. the effective weight equals v (torch.nn.utils.weight_norm at construction)
    c = gen.model["3"]
    torch.testing.assert_close(c.effective_weight(), c.weight_v.detach())
    gen.load_state_dict(synthetic.synthetic_melgan_state_dict(80, 32, 3, seed=3), strict=True)
    w = c.effective_weight()
    torch.testing.assert_close(w.flatten(1).norm(dim=1), c.weight_g.detach().flatten())   # |w| = g per input channel
    with pytest.raises(RuntimeError):
        gen(torch.zeros(1, 80, 16))          # CPU tensor: no fallback
    if os.path.isdir("/root/reference/vocoder"):
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        from make_golden_melgan import reference_generator
        ref = reference_generator(dict(n_mel=80, ngf=32, n_residual_layers=3), seed=3)
        assert set(ref.state_dict()) == set(sd)


def test_parameter_signature_cache_and_its_invalidation():
    """The wrappers decide whether libmgv's packed weight copies are stale from a signature over (data_ptr, version, device)
    of every parameter.  The tensor list behind it is cached (walking module.parameters() costs ~1.3 ms per call with the
    GPU idle); the cache must be rebuilt after _apply (.to / .double ...), load_state_dict and refresh_weights, in-place
    updates must change the signature, and copy.deepcopy / pickling must not carry the cache or the handle over."""
    import copy
    import pickle
    from melspec_gpt_vqvae_b200.transformer.minGPT import GPT, GPTConfig
    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE, _params_signature

    m = LitVQVAE(128, 256)
    sig = lambda: hash(tuple(_params_signature(x, m) for x in m._hot_modules()))
    s0 = sig()
    assert len(m.__dict__["_mgv_sig_tensors"]) == len(m._hot_modules()) and sig() == s0      # cached and stable
    n_cached = sum(len(v) for v in m.__dict__["_mgv_sig_tensors"].values())
    assert n_cached == sum(len(list(x.parameters())) + len(list(x.buffers())) for x in m._hot_modules())
    with torch.no_grad():
        m._decoder.conv_out.bias.add_(1.0)
    s1 = sig()
    assert s1 != s0, "in-place update not detected"
    m.double()                                                                                 # _apply: new storages
    assert m.__dict__["_mgv_sig_tensors"] == {}, "cache must be dropped by _apply"
    s2 = sig()
    assert s2 != s1
    m.load_state_dict(m.state_dict())
    assert m.__dict__["_mgv_sig_tensors"] == {}, "cache must be dropped by load_state_dict"
    assert sig() != s2                                                                          # copy_ bumps the versions
    sig()
    m.refresh_weights()
    assert m.__dict__["_mgv_sig_tensors"] == {} and m._mgv_sig is None
    m2 = copy.deepcopy(m)
    assert m2.__dict__["_mgv_sig_tensors"] == {} and m2._mgv_handle is None
    m3 = pickle.loads(pickle.dumps(m))
    assert m3.__dict__["_mgv_sig_tensors"] == {} and m3._mgv_handle is None

    g = GPT(GPTConfig(128, 265, n_layer=2, n_head=4, n_embd=64))
    a = _params_signature(g, g)
    assert _params_signature(g, g) == a and len(g.__dict__["_mgv_sig_tensors"]) == 1
    with torch.no_grad():
        g.head.weight.mul_(0.5)
    assert _params_signature(g, g) != a
    g.float()
    assert g.__dict__["_mgv_sig_tensors"] == {}
    assert copy.deepcopy(g).__dict__["_mgv_sig_tensors"] == {}


def test_extract_codes_loader_and_writer_host_logic(tmp_path):
    """Host stages of the batched extract_codes walk (no GPU): a batch is loaded straight into the staging buffer with the
    reference's `2 * crop(mel) - 1` (feature_extraction/extract_codes.py:40-43 of the reference), damaged files are reported
    and skipped, and the code files are byte-identical to what np.save writes (reference :58)."""
    import io
    from concurrent.futures import ThreadPoolExecutor
    from melspec_gpt_vqvae_b200.feature_extraction import extract_codes as ec
    d = tmp_path / "cls" / "melspec_10s_22050hz"
    d.mkdir(parents=True)
    rng = np.random.default_rng(1)
    paths = []
    for i in range(7):
        p = str(d / ("clip%02d_mel.npy" % i))
        arr = rng.random((80, 860)).astype(np.float64 if i == 2 else np.float32)       # one float64 file
        np.save(p, arr)
        paths.append(p)
    (d / "broken_mel.npy").write_bytes(b"not a numpy file")
    np.save(str(d / "short_mel.npy"), rng.random((80, 100), dtype=np.float32))          # narrower than the crop
    bad = [str(d / "broken_mel.npy"), str(d / "short_mel.npy")]
    tr = ec.Crop([80, 848], False)
    batch = paths[:3] + [bad[0]] + paths[3:5] + [bad[1]] + paths[5:]
    dst = np.full((len(batch), 80, 848), np.nan, dtype=np.float32)
    with ThreadPoolExecutor(max_workers=3) as pool:
        good, rows = ec._load_many(batch, tr, pool, dst, 3)
    assert good == paths and rows == [0, 1, 2, 4, 5, 7, 8]
    for p, j in zip(good, rows):
        ref = 2 * tr(np.load(p).astype(np.float32)) - 1
        assert np.array_equal(dst[j], ref), p
    codes = rng.integers(0, 128, (5, 53)).astype(np.int64)
    out = ec._out_path(paths[0], "codes_10s")
    assert out.endswith(os.path.join("cls", "codes_10s", "clip00_mel_code.npy"))
    ec._save_code(out, codes)
    ref = io.BytesIO()
    np.save(ref, codes)
    assert open(out, "rb").read() == ref.getvalue()
    ec._save_code(out, codes.astype(np.int32).T)                                         # another shape / dtype: its own header
    assert np.array_equal(np.load(out), codes.astype(np.int32).T)
