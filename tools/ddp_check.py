"""Data-parallel check of the minGPT training step (run under torchrun, one process per GPU):

  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py

(1) parity: the bucketed NCCL-averaged gradients of N ranks, each on its slice of a global batch, equal the gradients one
    process computes on the whole batch (reference: DDP wrapping, GPT_VAE_train.py:172-174);
(2) overlap: step time with the per-bucket all-reduce launched while the backward continues vs one all-reduce after the
    backward vs no all-reduce at all.  Prints one JSON line on rank 0."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from melspec_gpt_vqvae_b200 import synthetic
from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)


def make(cfg, pdrop):
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=False)
    args = argparse.Namespace(embd_pdrop=pdrop, resid_pdrop=pdrop, attn_pdrop=pdrop, reconstruct_spec="", device=dev,
                              learning_rate=1e-6, **cfg)
    lit = Lit_minGPT(args)
    lit.transformer.load_state_dict(sd, strict=False)
    return lit.to(dev).train()


out = {"world": world}
# ---- (1) parity on a 4-layer model of the VAS width, dropout off
cfg = dict(synthetic.GPT_VAS, n_layer=4, class_size=309)
per = 4
g = torch.Generator().manual_seed(11)
X = torch.randint(0, 128, (per * world, 265), generator=g)
Cc = torch.randint(0, 309, (per * world, 1), generator=g)
lit = make(cfg, 0.0)
tr = lit.trainer()
tr.layers_per_bucket = 2
lo = rank * per
loss = tr.step(X[lo:lo + per, :-1].to(dev), Cc[lo:lo + per].to(dev), X[lo:lo + per].to(dev))
tr.wait_for_gradients()
torch.cuda.synchronize()
g_ddp = tr.flat_grads.clone()
if rank == 0:
    ref = make(cfg, 0.0)
    rt = ref.trainer()
    rt.allreduce = False
    loss_ref = rt.step(X[:, :-1].to(dev), Cc.to(dev), X.to(dev))
    torch.cuda.synchronize()
    d = (g_ddp - rt.flat_grads).norm() / rt.flat_grads.norm()
    out["parity"] = {"grad_rel_l2_ddp_vs_single": float(d), "loss_ddp": float(loss), "loss_single": float(loss_ref),
                     "global_batch": per * world}
    del ref, rt
dist.barrier()
del lit, tr
torch.cuda.empty_cache()

# ---- (2) overlap on the full config-4 model, per-GPU batch 8, dropout 0.5
cfg = dict(synthetic.GPT_VAS, class_size=309)
lit = make(cfg, 0.5)
tr = lit.trainer()
opt = lit.configure_optimizers()
g = torch.Generator().manual_seed(100 + rank)
batch = {"codes": torch.randint(0, 128, (8, 5, 53), generator=g), "target": torch.randint(0, 309, (8,), generator=g)}


def timed(n=8):
    for it in range(3):
        lit.training_step(batch, it); opt.step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(n):
        lit.training_step(batch, it); opt.step()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


tr.allreduce = True; tr.overlap = True
ms_overlap = timed()
tr.overlap = False
ms_serial = timed()
tr.allreduce = False
ms_local = timed()
if rank == 0:
    comm = ms_serial - ms_local
    out["timing"] = {"ms_per_step_overlapped": ms_overlap, "ms_per_step_allreduce_after_backward": ms_serial,
                     "ms_per_step_no_allreduce": ms_local, "exposed_comm_ms_overlapped": ms_overlap - ms_local,
                     "comm_ms_serial": comm,
                     "overlap_fraction": (1.0 - (ms_overlap - ms_local) / comm) if comm > 1e-3 else None,
                     "allreduce_bytes": int(tr.flat_grads.numel()) * 4, "buckets": len(tr._buckets()),
                     "collective": "NCCL all-reduce (AVG) on torch's NCCL stream, one call per bucket of 4 blocks"}
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
