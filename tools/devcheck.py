"""Device diagnostics (not a test): cross-checks the tcgen05 GEMM / implicit-GEMM conv / VQ kernels
against torch on the GPU and prints error tables.  Run under gpurun; writes gpurun_out/devcheck.log."""
import ctypes, os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import _lib

L = _lib.load()
dev = torch.device("cuda:0")
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
log = open(os.path.join(out_dir, "devcheck.log"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n"); log.flush()


P("device", torch.cuda.get_device_name(0), "check", L.mgv_device_check())
s0 = ctypes.c_void_p(0)


def gemm(impl, A, B, epi, bias, out, resid, bn, split):
    M, K = A.shape; N = B.shape[0]
    rc = L.mgv_test_gemm(impl, _lib.ptr(A), _lib.ptr(B), M, N, K, epi, _lib.ptr(bias), _lib.ptr(out), _lib.ptr(resid), bn, split, s0)
    torch.cuda.synchronize()
    return rc


def check_gemm(M, N, K, bn, split=1, epi=2):
    torch.manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    ref = A.float() @ B.float().t() + bias
    resid = None
    if epi in (0, 1):
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        if epi == 1: ref = torch.nn.functional.gelu(ref)
    elif epi == 3:
        out = torch.randn(M, N, device=dev); resid = out; ref = ref + out.clone()
    elif epi == 4:
        out = torch.randn(M, N, device=dev); ref = ref + out.clone()
    elif epi == 5:
        resid = torch.randn(M, N, device=dev).bfloat16(); out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); ref = ref + resid.float()
    else:
        out = torch.zeros(M, N, device=dev)
    try:
        rc = gemm(0, A, B, epi, bias, out, resid, bn, split)
    except Exception as e:
        P("GEMM EXC", M, N, K, bn, split, epi, repr(e)); return False
    if rc != 0:
        P("GEMM rc", rc, _lib.last_error()); return False
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-2 * scale if epi in (0, 1, 5) else 1e-3 * scale
    ok = err <= tol
    P("gemm M=%d N=%d K=%d bn=%d split=%d epi=%d  maxerr=%.3e scale=%.3e  %s" % (M, N, K, bn, split, epi, err, scale, "OK" if ok else "FAIL"))
    if not ok:
        bad = ((out.float() - ref).abs() > tol).nonzero()
        P("   first bad idx:", bad[:8].tolist(), "n_bad", bad.shape[0], "of", M * N)
        P("   out[0,:8]", out[0, :8].float().tolist()); P("   ref[0,:8]", ref[0, :8].tolist())
    return ok


def check_conv(n, H, W, Cin, Cout, stride=1, resid=False):
    torch.manual_seed(H * W + Cin)
    x = (torch.randn(n, H, W, Cin, device=dev) * 0.5).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device=dev) * 0.05).bfloat16()
    bias = torch.randn(Cout, device=dev)
    xn = x.float().permute(0, 3, 1, 2); wn = w.float().permute(0, 3, 1, 2)
    if stride == 1:
        ref = torch.nn.functional.conv2d(xn, wn, bias, padding=1)
    else:
        ref = torch.nn.functional.conv2d(torch.nn.functional.pad(xn, (0, 1, 0, 1)), wn, bias, stride=2)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    Ho, Wo = ref.shape[1], ref.shape[2]
    r = None
    if resid:
        r = torch.randn(n, Ho, Wo, Cout, device=dev).bfloat16(); ref = ref + r.float()
    out = torch.zeros(n, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
    rc = L.mgv_test_conv3x3(0, _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), n, H, W, Cin, Cout, stride, _lib.ptr(out), _lib.ptr(r), s0)
    torch.cuda.synchronize()
    if rc != 0:
        P("CONV rc", rc, _lib.last_error()); return False
    err = (out.float() - ref).abs().max().item(); scale = ref.abs().max().item()
    ok = err <= 2e-2 * scale
    P("conv n=%d H=%d W=%d Cin=%d Cout=%d s=%d resid=%d maxerr=%.3e scale=%.3e %s" % (n, H, W, Cin, Cout, stride, resid, err, scale, "OK" if ok else "FAIL"))
    if not ok:
        bad = ((out.float() - ref).abs() > 2e-2 * scale).nonzero()
        P("   first bad idx:", bad[:8].tolist(), "n_bad", bad.shape[0], "of", out.numel())
    return ok


def check_vq(B, K=128, D=256, HW=265):
    torch.manual_seed(B)
    z = torch.randn(B, D, HW, device=dev) * 0.2
    cb = torch.randn(K, D, device=dev) * 0.2
    idx = torch.empty(B * HW, device=dev, dtype=torch.int64)
    dmin = torch.empty(B * HW, device=dev)
    rc = L.mgv_vq_argmin(_lib.ptr(z), _lib.ptr(cb), B, D, HW, K, _lib.ptr(idx), _lib.ptr(dmin), s0)
    torch.cuda.synchronize()
    if rc != 0:
        P("VQ rc", rc, _lib.last_error()); return False
    flat = z.permute(0, 2, 1).reshape(-1, D).double()
    d = (flat ** 2).sum(1, keepdim=True) + (cb.double() ** 2).sum(1) - 2 * flat @ cb.double().t()
    ref = d.argmin(1)
    mism = (ref != idx).sum().item()
    gap = (d.gather(1, idx[:, None]) - d.min(1, keepdim=True).values).abs().max().item()
    P("vq B=%d K=%d mismatches vs fp64 argmin: %d of %d, max dist gap at chosen idx %.3e" % (B, K, mism, idx.numel(), gap))
    return gap < 1e-4


def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


results = []
try:
    # smallest first: a hang / trap here tells us the basics are wrong
    results.append(check_gemm(128, 128, 64, 128))
    results.append(check_gemm(128, 128, 256, 128))
    results.append(check_gemm(128, 32, 1024, 32))
    results.append(check_gemm(64, 3072, 1024, 128))
    results.append(check_gemm(64, 3072, 1024, 128, split=8, epi=4))
    results.append(check_gemm(64, 1024, 4096, 64, split=8, epi=4))
    results.append(check_gemm(300, 256, 512, 256))
    results.append(check_gemm(1000, 384, 1024, 128, epi=0))
    results.append(check_gemm(1000, 4096, 1024, 128, epi=1))
    results.append(check_gemm(1000, 1024, 4096, 64, epi=3))
    results.append(check_gemm(530, 1472, 1472, 32, epi=5))
    results.append(check_vq(4)); results.append(check_vq(256)); results.append(check_vq(3, K=100, D=64, HW=17))
    results.append(check_conv(2, 5, 53, 256, 512))
    results.append(check_conv(2, 10, 106, 512, 512, resid=True))
    results.append(check_conv(1, 40, 424, 128, 128))
    results.append(check_conv(1, 80, 848, 128, 128))
    results.append(check_conv(2, 80, 848, 128, 128, stride=2))
    results.append(check_conv(2, 10, 106, 256, 256, stride=2))
    # timings
    for (M, N, K, bn, split, epi) in [(16960, 3072, 1024, 128, 1, 0), (16960, 1024, 4096, 128, 1, 3), (16960, 4096, 1024, 128, 1, 1),
                                      (64, 3072, 1024, 128, 8, 4), (64, 3072, 1024, 32, 1, 2), (64, 1024, 4096, 64, 8, 4), (64, 4096, 1024, 128, 4, 4),
                                      (64, 4096, 1024, 32, 1, 1), (64, 1024, 1024, 64, 8, 4)]:
        A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); bias = torch.randn(N, device=dev)
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
        resid = out if epi == 3 else None
        ms = timeit(lambda: L.mgv_test_gemm(0, _lib.ptr(A), _lib.ptr(B), M, N, K, epi, _lib.ptr(bias), _lib.ptr(out), _lib.ptr(resid), bn, split, s0))
        fl = 2.0 * M * N * K
        P("time gemm M=%d N=%d K=%d bn=%d split=%d epi=%d: %.4f ms  %.1f TFLOP/s  weights %.1f GB/s" % (M, N, K, bn, split, epi, ms, fl / ms / 1e9, N * K * 2 / ms / 1e6))
    for (n, H, W, Cin, Cout) in [(8, 80, 848, 128, 128), (8, 40, 424, 256, 256), (8, 5, 53, 512, 512)]:
        x = torch.randn(n, H, W, Cin, device=dev).bfloat16(); w = torch.randn(Cout, 3, 3, Cin, device=dev).bfloat16(); bias = torch.randn(Cout, device=dev)
        out = torch.zeros(n, H, W, Cout, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: L.mgv_test_conv3x3(0, _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), n, H, W, Cin, Cout, 1, _lib.ptr(out), None, s0))
        P("time conv n=%d %dx%d %d->%d: %.4f ms %.1f TFLOP/s" % (n, H, W, Cin, Cout, ms, 2.0 * n * H * W * Cin * Cout * 9 / ms / 1e9))
    z = torch.randn(256, 256, 265, device=dev) * 0.2; cb = torch.randn(128, 256, device=dev) * 0.2
    idx = torch.empty(256 * 265, device=dev, dtype=torch.int64)
    ms = timeit(lambda: L.mgv_vq_argmin(_lib.ptr(z), _lib.ptr(cb), 256, 256, 265, 128, _lib.ptr(idx), None, s0))
    P("time vq_argmin B=256: %.4f ms  %.1f GB/s algorithmic" % (ms, 256 * 265 * 1032 / ms / 1e6))
except Exception:
    P("EXCEPTION", traceback.format_exc())
P("SUMMARY", sum(1 for r in results if r), "ok of", len(results))
