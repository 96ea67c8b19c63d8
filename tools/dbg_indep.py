import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import make_vqvae, err_stats
from melspec_gpt_vqvae_b200 import synthetic
sd = synthetic.synthetic_vqvae_state_dict(128, 256, seed=783435, perturb=True)
m = make_vqvae(sd)
gen = torch.Generator().manual_seed(5)
codes = torch.randint(0, 128, (3, 265), generator=gen).cuda()
a = m.decode_codes(codes).cpu(); b = m.decode_codes(codes).cpu()
print("B=3 run-to-run max diff", float((a - b).abs().max()))
for i in range(3):
    o = m.decode_codes(codes[i:i+1]).cpu()
    print("clip", i, "B=1 vs B=3:", err_stats(o, a[i:i+1]))
o1 = m.decode_codes(codes[1:2]).cpu(); o2 = m.decode_codes(codes[1:2]).cpu()
print("B=1 run-to-run", float((o1 - o2).abs().max()))
c2 = torch.cat([codes[1:2], codes[1:2], codes[1:2]])
d = m.decode_codes(c2).cpu()
print("same clip x3: intra-batch diffs", float((d[0]-d[1]).abs().max()), float((d[0]-d[2]).abs().max()), "vs B=1", float((d[0]-o1[0]).abs().max()))
