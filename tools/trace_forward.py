"""CUPTI per-kernel totals of the teacher-forced forward (bs=64, T=265, VAS model).  Diagnostic tool (gpurun)."""
import argparse, collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT

dev = torch.device("cuda", 0)
cfg = synthetic.GPT_VAS
args = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, reconstruct_spec="", device=dev, **cfg)
lit = Lit_minGPT(args); lit.transformer.load_state_dict(synthetic.synthetic_gpt_state_dict(cfg, perturb=False), strict=False)
lit = lit.eval().to(dev)
x = torch.randint(0, 128, (64, 265)).to(dev); c = torch.randint(0, 8, (64, 1)).to(dev)
for _ in range(2): lit(x, c)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    lit(x, c); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, ""])
for e in prof.events():
    if e.device_type.name == "CUDA":
        m = re.search(r"([a-z_0-9]+_kernel(<[\d, a-z]+>)?)", e.name)
        k = m.group(1) if m else e.name[:50]
        agg[k][0] += 1; agg[k][1] += e.device_time
tot = sum(v[1] for v in agg.values())
for k, (n, t, _) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("%-44s n=%4d total %9.1f us (%4.1f%%) avg %8.1f" % (k, n, t, 100 * t / tot, t / n))
print("sum of kernels %.2f ms" % (tot / 1e3))
# the four GEMMs of the layers in launch order (QKV, proj, FC1 with the GELU epilogue, FC2): per-launch durations of layers 2..4
gem = [e for e in prof.events() if e.device_type.name == "CUDA" and "gemm_tc_persist" in e.name]
gem.sort(key=lambda e: e.time_range.start)
fl = [2 * 16960 * 3072 * 1024, 2 * 16960 * 1024 * 1024, 2 * 16960 * 4096 * 1024, 2 * 16960 * 1024 * 4096]
for l in range(2, 5):
    print("layer %d: " % l + "  ".join("%s %.1f us (%.0f TFLOP/s)" % (n, gem[4 * l + i].device_time, fl[i] / gem[4 * l + i].device_time / 1e6)
                                       for i, n in enumerate(("QKV", "proj", "FC1", "FC2"))))
