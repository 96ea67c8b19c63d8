"""CUPTI timeline of LitVQVAE.decode_codes / encode (B given): per-kernel totals.  Diagnostic tool."""
import collections, json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
what = sys.argv[2] if len(sys.argv) > 2 else "decode"
dev = torch.device("cuda", 0)
m = LitVQVAE(128, 256)
m.load_state_dict(synthetic.synthetic_vqvae_state_dict(128, 256, perturb=False), strict=False)
m = m.eval().to(dev)
codes = torch.randint(0, 128, (B, 265)).to(dev)
mel = torch.rand(B, 1, 80, 848, device=dev) * 2 - 1
fn = (lambda: m.decode_codes(codes)) if what == "decode" else (lambda: m.encode(mel))
for _ in range(2): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
print("%s B=%d: %.2f ms" % (what, B, e0.elapsed_time(e1)))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn(); torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "vq_trace.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
agg = collections.defaultdict(list)
for e in ev:
    mm = re.search(r"(gemm_tc_kernel<\d+>|[a-z_0-9]+_kernel(<\d+>)?)", e["name"])
    agg[mm.group(1) if mm else e["name"][:40]].append(e["dur"])
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-36s n=%3d total %9.1f us (%.1f%%)  max %8.1f" % (k, len(v), sum(v), 100 * sum(v) / tot, max(v)))
print("sum of kernels %.2f ms; span %.2f ms" % (tot / 1e3, (ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]) / 1e3))
big = sorted(ev, key=lambda e: -e["dur"])[:14]
for e in big:
    print("   %8.1f us grid=%s %s" % (e["dur"], e.get("args", {}).get("grid"), re.search(r"([a-z_0-9]+_kernel(<\d+>)?)", e["name"]).group(1)))
os.remove(out)
