"""GPU parity: fused quantiser kernels (through the drop-in VectorQuantizer -> C ABI) against the
oracle (bit-exact) and against the reference's golden outputs (indices equal except classified ties)."""
import numpy as np
import pytest
import torch

from helpers import golden
from make_golden import vq_inputs
from oracle import vq_oracle

pytestmark = pytest.mark.gpu


def _vq(K, D, cb, device="cuda"):
    from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import VectorQuantizer
    m = VectorQuantizer(K, D, 0.25)
    m._embedding.weight.data.copy_(cb)
    return m.to(device)


@pytest.mark.parametrize("case", ["trained", "default_init", "ties", "small"])
def test_vq_forward_parity(case):
    g = golden("vq_%s.npz" % case)
    z, cb = vq_inputs(case)
    K, D = cb.shape
    m = _vq(K, D, cb)
    loss, quant, (perp, enc, idx) = m(z.cuda())
    assert idx.dtype == torch.int64 and idx.shape == (z.shape[0] * z.shape[2] * z.shape[3], 1)
    assert quant.shape == z.shape and enc.shape == (idx.shape[0], K)
    idx_np = idx.cpu().numpy().reshape(-1)
    # (1) bit-exact against the oracle's fixed-order restatement (indices AND winning distances)
    o_idx, o_dmin = vq_oracle.argmin_exact(z.numpy(), cb.numpy())
    assert np.array_equal(idx_np, o_idx), "CUDA argmin differs from the oracle at %d rows" % int((idx_np != o_idx).sum())
    # (2) against the unmodified reference: equal except near-ties, which are counted and classified
    ref_idx = g["idx"].astype(np.int64).reshape(-1)
    dist = vq_oracle.distances_exact(z.numpy(), cb.numpy())
    rep = vq_oracle.classify_mismatches(dist, idx_np, ref_idx, ulps=16)
    print("vq %s vs reference: %s" % (case, rep))
    assert rep["n_real"] == 0, rep
    if case == "trained":
        assert rep["n_mismatch"] == 0
    # (3) rest of the tuple, evaluated on rows where indices agree with the reference
    same = idx_np == ref_idx
    q = quant.cpu().numpy()
    qf = np.transpose(q, (0, 2, 3, 1)).reshape(-1, D)
    rf = np.transpose(g["quantized"], (0, 2, 3, 1)).reshape(-1, D)
    np.testing.assert_allclose(qf[same], rf[same], rtol=0, atol=1e-7)
    if same.all():
        np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-6)
        np.testing.assert_allclose(perp.item(), g["perplexity"], rtol=1e-5)
        np.testing.assert_array_equal(enc.sum(0).cpu().numpy(), g["enc_colsum"])
    else:   # tie rows pick an equidistant code: loss/perplexity move by O(n_tie / N)
        np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-3)
    np.testing.assert_array_equal(enc.sum(1).cpu().numpy(), g["enc_rowsum"])
    assert torch.equal(enc.argmax(1), idx.reshape(-1))
    # full tuple against the numpy restatement on OUR indices (exact quantized, tight scalars)
    nl, nq, (npp, nenc, _) = vq_oracle.forward_numpy(z.numpy(), cb.numpy(), 0.25, indices=idx_np)
    np.testing.assert_array_equal(q, nq)
    np.testing.assert_array_equal(enc.cpu().numpy(), nenc)
    np.testing.assert_allclose(loss.item(), nl, rtol=1e-6)
    np.testing.assert_allclose(perp.item(), npp, rtol=1e-5)


@pytest.mark.parametrize("case", ["trained", "small"])
def test_get_codebook_entry_exact(case):
    g = golden("vq_%s.npz" % case)
    z, cb = vq_inputs(case)
    K, D = cb.shape
    m = _vq(K, D, cb)
    ref_idx = torch.from_numpy(g["idx"].astype(np.int64).reshape(-1)).cuda()
    out = m.get_codebook_entry(ref_idx, (z.shape[0], z.shape[2], z.shape[3], D))
    np.testing.assert_array_equal(out.cpu().numpy(), g["entry"])
    out2 = m.get_codebook_entry(ref_idx[:7], None)
    np.testing.assert_array_equal(out2.cpu().numpy(), g["entry_flat"])
    with pytest.raises(RuntimeError):
        m.get_codebook_entry(torch.tensor([0, K], device="cuda"), None)


def test_vq_full_size_config2():
    """BASELINE config 2 size: B=256 -> N=67 840 vectors; bit-exact against the oracle, plus the
    size-independent properties (idempotence: quantising codebook rows returns their own index)."""
    g = torch.Generator().manual_seed(0)
    z = torch.randn(256, 256, 5, 53, generator=g) * 0.2
    cb = torch.randn(128, 256, generator=g) * 0.2
    m = _vq(128, 256, cb)
    idx = m.encoding_indices(z.cuda()).cpu().numpy()
    o_idx, _ = vq_oracle.argmin_exact(z.numpy(), cb.numpy())
    assert np.array_equal(idx, o_idx)
    # idempotence: quantize(get_codebook_entry(idx)) == idx
    zq = m.get_codebook_entry(torch.from_numpy(idx).cuda(), (256, 5, 53, 256))
    idx2 = m.encoding_indices(zq).cpu().numpy()
    assert np.array_equal(idx2, idx)


def test_vq_edge_cases():
    g = torch.Generator().manual_seed(1)
    cb = torch.randn(128, 256, generator=g)
    m = _vq(128, 256, cb)
    # empty batch
    loss, quant, (perp, enc, idx) = m(torch.zeros(0, 256, 5, 53, device="cuda"))
    assert quant.shape == (0, 256, 5, 53) and idx.shape == (0, 1)
    # single vector, ragged spatial size (1x1), N not a multiple of the 128-vector tile
    for shape in [(1, 256, 1, 1), (3, 256, 7, 19), (1, 256, 5, 53)]:
        z = torch.randn(*shape, generator=g)
        idx = m.encoding_indices(z.cuda()).cpu().numpy()
        o_idx, _ = vq_oracle.argmin_exact(z.numpy(), cb.numpy())
        assert np.array_equal(idx, o_idx), shape
    # no CPU fallback
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 256, 5, 53))
    with pytest.raises(RuntimeError):
        m.cpu()(torch.zeros(1, 256, 5, 53, device="cuda"))


@pytest.mark.parametrize("K", [129, 300, 1024])
def test_vq_large_codebooks_multi_pass(K):
    """Codebooks beyond the 128-code register tiling (the reference's 1024-code VGGSound variant, README.md:182) run as
    passes over 128-code chunks: indices stay bit-exact, ties across chunks still go to the lowest index."""
    g = torch.Generator().manual_seed(K)
    cb = torch.randn(K, 256, generator=g) * 0.2
    cb[K - 1] = cb[3]            # duplicate code in the last chunk: index 3 must win
    cb[130 % K] = cb[5] if K > 130 else cb[130 % K]
    z = torch.randn(4, 256, 5, 53, generator=g) * 0.2
    z[0, :, 0, 0] = cb[3]        # a vector equal to the duplicated code
    m = _vq(K, 256, cb)
    loss, quant, (perp, enc, idx) = m(z.cuda())
    o_idx, _ = vq_oracle.argmin_exact(z.numpy(), cb.numpy())
    assert np.array_equal(idx.cpu().numpy().reshape(-1), o_idx)
    assert int(idx[0]) == 3
    r_loss, r_quant, (r_perp, r_enc, r_idx) = vq_oracle.forward_numpy(z.numpy(), cb.numpy(), 0.25, indices=o_idx)
    np.testing.assert_allclose(float(loss), float(r_loss), rtol=1e-5)
    np.testing.assert_allclose(float(perp), float(r_perp), rtol=1e-4)
    assert enc.shape == (4 * 265, K) and float(enc.sum()) == 4 * 265
    np.testing.assert_array_equal(quant.cpu().numpy(), r_quant)


def _argmin_abi(z, cb, want_dmin):
    """mgv_vq_argmin through ctypes with / without the exact-distance output (the two code paths of the tensor-core
    prefilter: with dmin every vector's winner is re-evaluated exactly; without it single-candidate vectors are final)."""
    from melspec_gpt_vqvae_b200 import _lib
    B, D, H, W = z.shape
    K = cb.shape[0]
    zc, cc = z.cuda().contiguous(), cb.cuda().contiguous()
    idx = torch.full((B * H * W,), -1, dtype=torch.int64, device="cuda")
    dmin = torch.full((B * H * W,), float("nan"), dtype=torch.float32, device="cuda") if want_dmin else None
    _lib.check(_lib.load().mgv_vq_argmin(_lib.ptr(zc), _lib.ptr(cc), B, D, H * W, K, _lib.ptr(idx), _lib.ptr(dmin),
                                         _lib.stream_ptr(zc.device)), "mgv_vq_argmin")
    torch.cuda.synchronize()
    return idx.cpu().numpy(), (dmin.cpu().numpy() if want_dmin else None)


@pytest.mark.parametrize("name", ["gauss", "all_codes_equal", "zero_vectors", "vectors_on_codes", "near_ties", "k5_d64",
                                  "large_dynamic_range"])
def test_vq_prefilter_candidates_are_exact(name):
    """The TF32 prefilter may only narrow the search: indices AND winning distances must stay bit-identical to the
    oracle's exhaustive fixed-order evaluation, including inputs built to defeat an approximate search."""
    g = torch.Generator().manual_seed(len(name))
    K, D = 128, 256
    cb = torch.randn(K, D, generator=g) * 0.2
    z = torch.randn(5, D, 5, 53, generator=g) * 0.2
    if name == "all_codes_equal":          # 128 candidates per vector: index 0 must win everywhere
        cb = cb[:1].repeat(K, 1)
    elif name == "zero_vectors":
        z[1:3] = 0.0
    elif name == "vectors_on_codes":       # d = 0 for one code, duplicates of it later in the codebook
        cb[100] = cb[7]
        z = cb[torch.randint(0, K, (5 * 265,), generator=g)].reshape(5, 5, 53, D).permute(0, 3, 1, 2).contiguous()
    elif name == "near_ties":              # codes that differ from each other in the last bits only
        base = torch.randn(D, generator=g) * 0.2
        cb = base[None, :] * (1.0 + torch.arange(K)[:, None].float() * 2.0 ** -22)
    elif name == "k5_d64":
        K, D = 5, 64
        cb = torch.randn(K, D, generator=g)
        z = torch.randn(3, D, 7, 19, generator=g)
    elif name == "large_dynamic_range":
        z = z * torch.logspace(-3, 3, 5)[:, None, None, None]
        cb = cb * torch.logspace(-2, 2, K)[:, None]
    o_idx, o_dmin = vq_oracle.argmin_exact(z.numpy(), cb.numpy())
    i1, d1 = _argmin_abi(z, cb, True)
    i0, _ = _argmin_abi(z, cb, False)
    assert np.array_equal(i1, o_idx), "exact-distance path: %d rows differ" % int((i1 != o_idx).sum())
    assert np.array_equal(i0, o_idx), "direct path: %d rows differ" % int((i0 != o_idx).sum())
    assert np.array_equal(d1.view(np.uint32), o_dmin.view(np.uint32)), "winning distances are not bit-identical"
    if name == "all_codes_equal":
        assert (i0 == 0).all()
