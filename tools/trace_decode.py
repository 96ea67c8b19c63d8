"""CUPTI timeline (torch.profiler) of the decode loop: per-kernel durations and inter-kernel gaps inside
the CUDA-graph replay, without ncu's serialisation.  Diagnostic tool (run under gpurun)."""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT

dev = torch.device("cuda", 0)
cfg = synthetic.GPT_VAS
sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=False)
args = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, reconstruct_spec="", device=dev, **cfg)
lit = Lit_minGPT(args); lit.transformer.load_state_dict(sd, strict=False); lit = lit.eval().to(dev); lit.return_attention = False
t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
c = torch.randint(0, 8, (64, 1)).to(dev); x0 = torch.randint(0, 128, (64, t0)).to(dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
lit.sample(x0, c, steps=steps, sample=True, top_k=100)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    lit.sample(x0, c, steps=steps, sample=True, top_k=100)
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "decode_trace.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
import re
def kname(e):
    m = re.search(r"(gemm_tc_kernel<\d+>|[a-z_0-9]+_kernel)", e["name"])
    n = m.group(1) if m else e["name"][:30]
    g = e.get("args", {}).get("grid", "")
    return "%s grid=%s stream=%s" % (n, g, e.get("args", {}).get("stream", ""))
print("kernels traced:", len(ev))
dur = collections.defaultdict(list); gap = collections.defaultdict(list)
for i, e in enumerate(ev):
    name = kname(e).split(" stream=")[0]
    dur[name].append(e["dur"])
    if i > 0:
        gap[name].append(e["ts"] - (ev[i - 1]["ts"] + ev[i - 1]["dur"]))
tot = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
print("span %.1f us over %d positions = %.1f us/position" % (tot, steps, tot / steps))
for k in dur:
    g = gap[k]
    print("%-42s n=%5d dur avg %.2f us (sum %.0f)  gap-before avg %.2f us (sum %.0f)" % (k, len(dur[k]), sum(dur[k]) / len(dur[k]), sum(dur[k]), sum(g) / max(len(g), 1), sum(g)))
busy = 0.0; cur_s = None; cur_e = None
for e in ev:
    a, b = e["ts"], e["ts"] + e["dur"]
    if cur_e is None or a > cur_e:
        if cur_e is not None: busy += cur_e - cur_s
        cur_s, cur_e = a, b
    else:
        cur_e = max(cur_e, b)
busy += cur_e - cur_s
print("union of kernel intervals %.0f us of span %.0f us ; sum of durations %.0f us" % (busy, tot, sum(sum(v) for v in dur.values())))
# stage cost = time from the previous kernel's END to this kernel's END (the chain is serial), averaged per kernel
stage = collections.defaultdict(list)
for i in range(1, len(ev)):
    stage[kname(ev[i]).split(" stream=")[0]].append(ev[i]["ts"] + ev[i]["dur"] - (ev[i - 1]["ts"] + ev[i - 1]["dur"]))
print("stage cost (end-to-end delta to the previous kernel), average us and share of the span:")
for k, v in stage.items():
    print("  %-42s n=%5d avg %6.2f us  sum %8.0f us  %5.1f %%" % (k, len(v), sum(v) / len(v), sum(v), 100.0 * sum(v) / tot))
# one layer of the last position in detail
last = ev[-60:-30]
t0 = last[0]["ts"]
for e in last:
    print("  +%8.2f us  dur %6.2f  end +%8.2f  %s" % (e["ts"] - t0, e["dur"], e["ts"] + e["dur"] - t0, kname(e)))
os.remove(out)
