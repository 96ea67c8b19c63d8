"""Prefill attention alone (tcgen05 kernel vs the mma.sync kernel): parity against torch SDPA, CUDA-event timings, and the
clock stamps of one CTA of the tcgen05 kernel (phases of the TMA / MMA thread and of two softmax threads).
Diagnostic tool (run under gpurun):  python tools/attn_trace.py [B] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 265
nh, C = 16, 1024
L = _lib.load()
S0 = _lib.stream_ptr(torch.device("cuda", 0))
torch.manual_seed(0)
qkv = (torch.randn(B * T, 3 * C, device="cuda") * 1.0).bfloat16()
q, k, v = [t.float().view(B, T, nh, 64).transpose(1, 2) for t in qkv.split(C, dim=1)]
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B * T, C)
for impl in (0, 1):
    y = torch.zeros(B * T, C, device="cuda", dtype=torch.bfloat16)
    trace = torch.zeros(48, device="cuda", dtype=torch.int64)
    _lib.check(L.mgv_test_attention_prefill(impl, _lib.ptr(qkv), B, T, nh, _lib.ptr(y), _lib.ptr(trace) if impl == 0 else None, S0))
    torch.cuda.synchronize()
    err = (y.float() - ref).abs().max().item()
    y.zero_()
    L.mgv_test_attention_prefill(impl, _lib.ptr(qkv), B, T, nh, _lib.ptr(y), None, S0)   # untraced: the persistent form where it applies
    torch.cuda.synchronize()
    err = max(err, (y.float() - ref).abs().max().item())
    for _ in range(3):
        L.mgv_test_attention_prefill(impl, _lib.ptr(qkv), B, T, nh, _lib.ptr(y), None, S0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.mgv_test_attention_prefill(impl, _lib.ptr(qkv), B, T, nh, _lib.ptr(y), None, S0)
    e1.record(); torch.cuda.synchronize()
    print("impl %d (%s): max |err| vs fp32 SDPA %.4f, %.1f us per call" % (impl, "tcgen05" if impl == 0 else "mma.sync", err, e0.elapsed_time(e1) * 50))
    if impl == 0:
        t = trace.cpu().tolist()
        for name, off in (("thread 0 (TMA/MMA)", 0), ("thread 96 (quarter 3, half 0)", 16), ("thread 224 (quarter 3, half 1)", 32)):
            st = [x for x in t[off:off + 16] if x]
            print("  %-32s" % name, " ".join("%6d" % (x - t[0]) for x in st))
