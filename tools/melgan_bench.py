"""MelGAN vocoder timing: Generator(80, 32, 3) on B clips of 848 mel frames (217 088 samples
each), CUDA events, against the fp32-FMA roofline (77.2 GFLOP per clip) and against the reference algorithm in torch
eager fp32 on the same GPU (the oracle's functional form, cuDNN convolutions).  Diagnostic tool (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.vocoder.modules import Generator
from oracle import melgan_oracle

FLOP_PER_CLIP = 77.2e9
cfg = dict(n_mel=80, ngf=32, n_residual_layers=3)
sd = synthetic.synthetic_melgan_state_dict(seed=7, **cfg)
gen = Generator(80, 32, 3); gen.load_state_dict(sd); gen = gen.eval().cuda()
sd_dev = {k: v.cuda() for k, v in sd.items()}
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for B in (1, 8, 64):
    mel = torch.rand(B, 80, 848, device="cuda")
    def t(fn, n):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms = t(lambda: gen(mel), 5 if B < 64 else 2)
    with torch.no_grad():
        ms_ref = t(lambda: melgan_oracle.generator_forward(sd_dev, mel), 3 if B < 64 else 1)
        err = float((gen(mel) - melgan_oracle.generator_forward(sd_dev, mel)).abs().max())
    print("melgan B=%d: %.2f ms (%.0f clips/s, %.1f TFLOP/s fp32 = %.2f of the 74.5 TFLOP/s FMA peak; %d launches) | torch eager fp32 "
          "(TF32 off) %.2f ms -> x%.2f | max abs diff %.1e" % (B, ms, B / ms * 1e3, FLOP_PER_CLIP * B / ms / 1e9, FLOP_PER_CLIP * B / ms / 1e9 / 74.5,
                                                              gen.last_launches(), ms_ref, ms_ref / ms, err))
