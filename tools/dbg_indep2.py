import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from melspec_gpt_vqvae_b200 import _lib
L = _lib.load(); S0 = ctypes.c_void_p(0); dev = "cuda"
torch.manual_seed(0)
for (H, W, Cin, Cout) in [(5, 53, 256, 512), (5, 53, 512, 512), (10, 106, 512, 256), (20, 212, 256, 256)]:
    x1 = (torch.randn(1, H, W, Cin, device=dev) * 0.5).bfloat16()
    x = x1.repeat(3, 1, 1, 1).contiguous()
    w = (torch.randn(Cout, 3, 3, Cin, device=dev) * 0.05).bfloat16(); bias = torch.randn(Cout, device=dev)
    r1 = torch.randn(1, H, W, Cout, device=dev).bfloat16(); r = r1.repeat(3, 1, 1, 1).contiguous()
    for it in range(3):
        out = torch.zeros(3, H, W, Cout, device=dev, dtype=torch.bfloat16)
        _lib.check(L.mgv_test_conv3x3(0, _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), 3, H, W, Cin, Cout, 1, _lib.ptr(out), _lib.ptr(r), S0))
        torch.cuda.synchronize()
        o1 = torch.zeros(1, H, W, Cout, device=dev, dtype=torch.bfloat16)
        _lib.check(L.mgv_test_conv3x3(0, _lib.ptr(x1), _lib.ptr(w), _lib.ptr(bias), 1, H, W, Cin, Cout, 1, _lib.ptr(o1), _lib.ptr(r1), S0))
        torch.cuda.synchronize()
        print("conv %dx%d %d->%d it%d: img1-img0 %.4f img2-img0 %.4f  B1-img0 %.4f" % (H, W, Cin, Cout, it, float((out[1].float()-out[0].float()).abs().max()), float((out[2].float()-out[0].float()).abs().max()), float((o1[0].float()-out[0].float()).abs().max())))
# plain gemm: 3 identical row blocks of 265 rows
A1 = (torch.randn(265, 512, device=dev) * 0.5).bfloat16(); A = A1.repeat(3, 1).contiguous()
Wt = (torch.randn(1536, 512, device=dev) * 0.05).bfloat16(); bias = torch.randn(1536, device=dev)
for it in range(3):
    out = torch.zeros(795, 1536, device=dev, dtype=torch.bfloat16)
    _lib.check(L.mgv_test_gemm(0, _lib.ptr(A), _lib.ptr(Wt), 795, 1536, 512, 0, _lib.ptr(bias), _lib.ptr(out), None, 128, 1, S0)); torch.cuda.synchronize()
    o = out.view(3, 265, 1536).float()
    print("gemm it%d: blk1-blk0 %.4f blk2-blk0 %.4f" % (it, float((o[1]-o[0]).abs().max()), float((o[2]-o[0]).abs().max())))
