"""Quantiser alone at BASELINE config 2's size (256 clips -> N = 67 840 vectors, 128 codes x 256 channels): CUDA-event
time per call for the index-only path and the exact-distance path, per-kernel CUPTI durations, and the candidate
statistics of the prefilter.  Diagnostic tool (run under gpurun)."""
import collections, json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from melspec_gpt_vqvae_b200 import _lib

HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(0)
z = (torch.randn(B, 256, 5, 53, generator=g) * 0.2).cuda()
cb = (torch.randn(128, 256, generator=g) * 0.2).cuda()
N = B * 265
idx = torch.empty(N, dtype=torch.int64, device="cuda")
dmin = torch.empty(N, dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L = _lib.load()


def call(want_dmin):
    _lib.check(L.mgv_vq_argmin(_lib.ptr(z), _lib.ptr(cb), B, 256, 265, 128, _lib.ptr(idx), _lib.ptr(dmin) if want_dmin else None,
                               _lib.stream_ptr(z.device)), "mgv_vq_argmin")


for want in (False, True):
    for _ in range(3):
        call(want)
    ts = []
    for _ in range(20):
        if not os.environ.get('NOFLUSH'): flush.zero_()                      # z (69.5 MB) must come from HBM, not from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(want); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    alg = N * 1032
    print("quantiser N=%d %s: median %.1f us (min %.1f)  %.0f GB/s algorithmic = %.3f of HBM peak" %
          (N, "idx+dmin" if want else "idx only", us, ts[0], alg / us / 1e3, alg / us / 1e3 / HBM))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            (None if os.environ.get('NOFLUSH') else flush.zero_()); call(want)
        torch.cuda.synchronize()
    d = collections.defaultdict(list)
    for e in prof.events():
        m = re.search(r"vq_\w+", e.name)
        if m:
            d[m.group(0)].append(e.device_time)
    for k, v in d.items():
        print("    %-40s n=%d avg %.1f us" % (k, len(v), sum(v) / len(v)))
