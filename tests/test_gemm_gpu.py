"""GPU: the tcgen05 GEMM / implicit-GEMM conv kernel against the SIMT fp32 reference kernel of the same
contract (exactly the same bf16 inputs, fp32 accumulation -> tight tolerance) and against torch."""
import ctypes

import pytest
import torch

from melspec_gpt_vqvae_b200 import _lib

pytestmark = pytest.mark.gpu
S0 = ctypes.c_void_p(0)


def run_gemm(impl, A, B, epi, bias, out, resid, bn, split):
    M, K = A.shape
    N = B.shape[0]
    _lib.check(_lib.load().mgv_test_gemm(impl, _lib.ptr(A), _lib.ptr(B), M, N, K, epi, _lib.ptr(bias), _lib.ptr(out),
                                         _lib.ptr(resid), bn, split, S0))
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,bn,split,epi", [
    (128, 128, 64, 128, 1, 2), (1, 32, 64, 32, 1, 2), (64, 3072, 1024, 128, 8, 4), (64, 1024, 4096, 64, 8, 4),
    (64, 4096, 1024, 128, 4, 4), (64, 128, 1024, 128, 1, 2), (300, 256, 512, 256, 1, 2), (1000, 384, 1024, 128, 1, 0),
    (777, 4096, 1024, 128, 1, 1), (1000, 1024, 4096, 64, 1, 3), (530, 1472, 1472, 32, 1, 5), (16960, 1024, 1024, 128, 1, 3),
])
def test_gemm_vs_simt_reference(M, N, K, bn, split, epi):
    torch.manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    bf16_out = epi in (0, 1, 5)
    init = torch.randn(M, N, device="cuda")
    outs = []
    for impl in (0, 1):
        out = (torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if bf16_out else init.clone())
        resid = None
        if epi == 3:
            resid = out
        if epi == 5:
            resid = init.bfloat16()
        run_gemm(impl, A, B, epi, bias, out, resid, bn, split if impl == 0 else 1)
        outs.append(out.float())
    ref = A.float() @ B.float().t() + bias
    scale = ref.abs().max().item()
    err = (outs[0] - outs[1]).abs().max().item()
    tol = (1.0 / 128) * scale if bf16_out else 2e-5 * scale * (K ** 0.5)
    assert err <= tol, "tcgen05 vs SIMT: max err %.3e (tol %.3e)" % (err, tol)
    if epi == 2:
        assert (outs[0] - ref).abs().max().item() <= 1e-3 * scale


@pytest.mark.parametrize("M,N,K,bn,split,epi", [
    (3072, 64, 1024, 64, 8, 4), (1024, 64, 1024, 64, 16, 4), (4096, 64, 1024, 64, 4, 4), (1024, 64, 4096, 64, 16, 4),
    (3072, 4, 1024, 32, 8, 4), (1024, 1, 1024, 32, 1, 3), (4096, 37, 1024, 64, 1, 1), (128, 200, 256, 256, 1, 2),
    (1472, 64, 1472, 64, 1, 2),
])
def test_gemm_swapab_vs_torch(M, N, K, bn, split, epi):
    """decode-step form: weights (M,K) are the 128-row operand, N batch rows the MMA N dim, output transposed."""
    torch.manual_seed(M + N + K)
    W = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    X = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    bias = torch.randn(M, device="cuda")
    ref = X.float() @ W.float().t() + bias          # (N, M)
    init = torch.randn(N, M, device="cuda")
    if epi == 1:
        out = torch.zeros(N, M, device="cuda", dtype=torch.bfloat16)
        ref = torch.nn.functional.gelu(ref)
    else:
        out = init.clone()
        if epi in (3, 4):
            ref = ref + init
    resid = out if epi == 3 else None
    _lib.check(_lib.load().mgv_test_gemm_swapab(0, _lib.ptr(W), _lib.ptr(X), M, N, K, epi, _lib.ptr(bias), _lib.ptr(out),
                                                _lib.ptr(resid), bn, split, S0))
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    tol = (1.0 / 128) * scale if epi == 1 else 1e-3 * scale
    assert (out.float() - ref).abs().max().item() <= tol


@pytest.mark.parametrize("Nw,B,K,kbps,sw", [
    (3072, 64, 1024, 4, 4), (4096, 64, 1024, 4, 8), (128, 64, 1024, 1, 4), (3072, 3, 1024, 4, 4), (256, 64, 128, 2, 4),
    (576, 37, 192, 3, 8), (1024, 130, 1024, 2, 4), (3072, 64, 1024, 4, 108), (1024, 130, 1024, 4, 108), (576, 37, 192, 3, 108),
])
def test_fold_ln_decode_gemm(Nw, B, K, kbps, sw):
    """FOLD_LN: W LN(x) computed as rstd * (W bf16(gamma x) - mu W gamma) + W beta + b (decode chain, gemm_decode_fold.cu)
    against LayerNorm + matmul in fp32 (reference: Block.forward transformer/minGPT.py:107-119).  The rows carry a
    non-zero mean so that the mean-cancellation term is exercised."""
    torch.manual_seed(Nw + B + K)
    W = (torch.randn(Nw, K, device="cuda") * 0.05).bfloat16()
    src = torch.randn(B, K, device="cuda") * 2.0 + 0.3
    gamma = 1 + 0.1 * torch.randn(K, device="cuda")
    beta = 0.1 * torch.randn(K, device="cuda")
    bias = torch.randn(Nw, device="cuda")
    out = torch.zeros(B, Nw, device="cuda")
    ref = torch.nn.functional.layer_norm(src, (K,), gamma, beta, 1e-5) @ W.float().t() + bias
    _lib.check(_lib.load().mgv_test_gemm_fold(0, _lib.ptr(W), _lib.ptr(src), Nw, B, K, _lib.ptr(gamma), _lib.ptr(beta),
                                              _lib.ptr(bias), None, 0, 0, _lib.ptr(out), kbps, sw, S0))
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item()
    assert err <= 4e-3 * scale, "max err %.3e (scale %.3e)" % (err, scale)


@pytest.mark.parametrize("Nw,B,K,kbps,sw,nparts", [
    (1024, 64, 4096, 4, 8, 4), (1024, 64, 4096, 4, 4, 4), (1024, 5, 4096, 4, 8, 16), (128, 64, 256, 4, 8, 1),
    (192, 37, 768, 3, 4, 4), (1024, 64, 4096, 4, 108, 4), (192, 37, 768, 3, 108, 4), (1024, 130, 4096, 4, 108, 4),
])
def test_fold_gelu_decode_gemm(Nw, B, K, kbps, sw, nparts):
    """FOLD_GELU: FC2 applies the producer's folded LayerNorm, the bias and the erf GELU while it stages its operand
    (reference: mlp transformer/minGPT.py:100-105)."""
    torch.manual_seed(Nw + B + K + nparts)
    dim = 1024
    W = (torch.randn(Nw, K, device="cuda") * 0.05).bfloat16()
    src = torch.randn(B, K, device="cuda") * 3.0
    swv = torch.randn(K, device="cuda")
    bpv = 0.2 * torch.randn(K, device="cuda")
    bias = torch.randn(Nw, device="cuda")
    x = torch.randn(B, dim, device="cuda") * 1.5 + 0.4            # the LayerNorm input whose statistics travel
    parts = x.view(B, nparts, dim // nparts)
    stats = torch.stack([parts.sum(-1), (parts * parts).sum(-1)], dim=-1).permute(1, 0, 2).contiguous()   # (nparts, B, 2)
    mu = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
    act = torch.nn.functional.gelu(rstd * (src - mu * swv) + bpv)
    init = torch.randn(B, Nw, device="cuda")
    out = init.clone()
    ref = init + act.bfloat16().float() @ W.float().t() + bias
    _lib.check(_lib.load().mgv_test_gemm_fold(1, _lib.ptr(W), _lib.ptr(src), Nw, B, K, _lib.ptr(swv), _lib.ptr(bpv),
                                              _lib.ptr(bias), _lib.ptr(stats), nparts, dim, _lib.ptr(out), kbps, sw, S0))
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item()
    assert err <= 4e-3 * scale, "max err %.3e (scale %.3e)" % (err, scale)


@pytest.mark.parametrize("n,H,W,Cin,Cout,stride,resid", [
    (2, 5, 53, 256, 512, 1, False), (2, 10, 106, 512, 512, 1, True), (1, 40, 424, 128, 128, 1, False),
    (1, 80, 848, 128, 128, 1, True), (2, 80, 848, 128, 128, 2, False), (2, 10, 106, 256, 256, 2, False),
    (3, 20, 212, 256, 128, 1, False),
])
def test_conv3x3_vs_torch(n, H, W, Cin, Cout, stride, resid):
    torch.manual_seed(H * W + Cin)
    x = (torch.randn(n, H, W, Cin, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(Cout, device="cuda")
    xn, wn = x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2)
    if stride == 1:
        ref = torch.nn.functional.conv2d(xn, wn, bias, padding=1)
    else:   # Downsample: pad (0,1,0,1) then stride 2 (reference big_model_attn_gan.py:151-159)
        ref = torch.nn.functional.conv2d(torch.nn.functional.pad(xn, (0, 1, 0, 1)), wn, bias, stride=2)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    r = None
    if resid:
        r = torch.randn_like(ref).bfloat16()
        ref = ref + r.float()
    out = torch.zeros(ref.shape, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.load().mgv_test_conv3x3(0, _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), n, H, W, Cin, Cout, stride,
                                            _lib.ptr(out), _lib.ptr(r), S0))
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (out.float() - ref).abs().max().item() <= (1.0 / 64) * scale


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("n,H,W,Cin,Cout", [
    (2, 5, 53, 512, 512), (3, 10, 106, 256, 256), (1, 20, 212, 256, 256), (2, 40, 424, 128, 128), (1, 7, 19, 64, 96),
])
def test_conv_upsample_phase_form_vs_torch(impl, n, H, W, Cin, Cout):
    """Upsample.forward (reference big_model_attn_gan.py:182-186): nearest 2x + conv3x3 pad 1, computed as four 2x2
    convolutions on the low-res tensor with pre-summed weights, against torch on the materialised upsampled tensor."""
    if impl == 1 and H * W > 3000:
        pytest.skip("SIMT reference: small cases only")
    torch.manual_seed(H * W + Cin)
    x = (torch.randn(n, H, W, Cin, device="cuda") * 0.5).bfloat16()
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    bias = torch.randn(Cout, device="cuda")
    up = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = torch.nn.functional.conv2d(up, w, bias, padding=1).permute(0, 2, 3, 1).contiguous()
    out = torch.full(ref.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
    scratch = torch.empty(16 * Cout * Cin, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.load().mgv_test_conv_upsample(impl, _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), n, H, W, Cin, Cout,
                                                  _lib.ptr(out), _lib.ptr(scratch), S0))
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), "every output pixel of every phase must be written"
    scale = ref.abs().max().item()
    assert (out.float() - ref).abs().max().item() <= (1.0 / 64) * scale


def test_gemm_rejects_bad_shapes():
    A = torch.zeros(4, 60, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(32, 60, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(4, 32, device="cuda")
    rc = _lib.load().mgv_test_gemm(0, _lib.ptr(A), _lib.ptr(B), 4, 32, 60, 2, None, _lib.ptr(out), None, 32, 1, S0)
    assert rc != 0 and "multiple of 64" in _lib.last_error()
