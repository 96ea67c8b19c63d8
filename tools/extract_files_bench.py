"""the ON-DISK extract_codes path.  Writes N synthetic `*_mel.npy` files (80 x 860 fp32, the
reference's feature format) into a scratch directory, runs the drop-in's batched walk (thread-pool loader, pinned
double buffer, asynchronous copies, code files written by the pool) and reports clips/s end to end -- files in, files
out -- next to the in-memory encode + quantise rate of the same batch size.  Diagnostic tool (run under gpurun)."""
import json, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.feature_extraction import extract_codes as ec
from melspec_gpt_vqvae_b200.vqvae.big_model_attn_gan import LitVQVAE

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
BS = int(sys.argv[2]) if len(sys.argv) > 2 else 256
THREADS = int(sys.argv[3]) if len(sys.argv) > 3 else min(32, os.cpu_count() or 8)
dev = torch.device("cuda", 0)
model = LitVQVAE(128, 256)
model.load_state_dict(synthetic.synthetic_vqvae_state_dict(128, 256, perturb=True, codebook_scale=0.05), strict=False)
model = model.eval().to(dev)
root = tempfile.mkdtemp(prefix="mgv_extract_")
try:
    feat = os.path.join(root, "melspec_10s_22050hz")
    os.makedirs(feat)
    rng = np.random.default_rng(0)
    base = (rng.random((80, 860), dtype=np.float32) * 2 - 1)
    t0 = time.time()
    for i in range(N):
        np.save(os.path.join(feat, "clip_%06d_mel.npy" % i), np.roll(base, i, axis=1))
    print("wrote %d mel files (%.1f MB) in %.1f s" % (N, N * 80 * 860 * 4 / 1e6, time.time() - t0))
    paths = sorted(os.path.join(feat, f) for f in os.listdir(feat))
    transforms = ec.Crop([80, 848], False)
    # warm-up on a few files (library load, weight packing), into a throw-away folder name
    ec.get_codes_batch(paths[:BS], dev, 848, model, transforms, folder_name="codes_warm", batch_size=BS, io_threads=THREADS)
    torch.cuda.synchronize()
    t0 = time.time()
    done = ec.get_codes_batch(paths, dev, 848, model, transforms, folder_name="codes_10s", batch_size=BS, io_threads=THREADS)
    torch.cuda.synchronize()
    dt = time.time() - t0
    out = sorted(os.listdir(os.path.join(root, "codes_10s")))
    assert done == N and len(out) == N
    c0 = np.load(os.path.join(root, "codes_10s", out[0]))
    assert c0.shape == (5, 53) and c0.dtype == np.int64
    # in-memory rate of the same batch size
    mel = torch.rand(BS, 1, 80, 848, device=dev) * 2 - 1
    for _ in range(2):
        model._vq_vae.encoding_indices(model.encode(mel))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model._vq_vae.encoding_indices(model.encode(mel))
    e1.record(); torch.cuda.synchronize()
    mem = BS * 3 / (e0.elapsed_time(e1) / 1e3)
    res = {"files": N, "batch_size": BS, "io_threads": THREADS, "host_cores": os.cpu_count(), "files_per_s": N / dt,
           "seconds": dt, "in_memory_clips_per_s": mem, "file_path_fraction_of_in_memory": (N / dt) / mem,
           "scratch": "tmpfs/overlay of the GPU box (no real disk latency)"}
    print(json.dumps(res))
finally:
    shutil.rmtree(root, ignore_errors=True)
