"""CUPTI timeline of one minGPT training step (config 4 size): kernel time by name and launch gaps.  Diagnostic tool (gpurun)."""
import argparse, collections, json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(synthetic.GPT_VAS, class_size=309)
sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=False)
args = argparse.Namespace(embd_pdrop=0.5, resid_pdrop=0.5, attn_pdrop=0.5, reconstruct_spec="", device="cuda", learning_rate=1e-6, **cfg)
lit = Lit_minGPT(args); lit.transformer.load_state_dict(sd, strict=False); lit = lit.to("cuda").train()
opt = lit.configure_optimizers()
g = torch.Generator().manual_seed(8)
batch = {"codes": torch.randint(0, 128, (B, 5, 53), generator=g), "target": torch.randint(0, 309, (B,), generator=g)}
for it in range(3):
    lit.training_step(batch, it); opt.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(5):
    lit.training_step(batch, it); opt.step()
e1.record(); torch.cuda.synchronize()
print("B=%d: %.2f ms per training step (events)" % (B, e0.elapsed_time(e1) / 5))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    lit.training_step(batch, 0); opt.step()
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "train_trace.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
dur = collections.defaultdict(float); cnt = collections.Counter()
for e in ev:
    m = re.search(r"([a-z_0-9]+_kernel(<[^>]*>)?)", e["name"])
    n = (m.group(1) if m else e["name"])[:60]
    dur[n] += e["dur"]; cnt[n] += 1
tot = sum(dur.values())
print("span %.0f us, kernel time %.0f us (%.0f %% busy), %d launches" % (span, tot, 100 * tot / span, len(ev)))
for n, d in sorted(dur.items(), key=lambda kv: -kv[1])[:25]:
    print("  %-62s n=%4d  %8.0f us  %5.1f %%" % (n, cnt[n], d, 100 * d / tot))
os.remove(out)
