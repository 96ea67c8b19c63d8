"""Training step of the class-conditioned minGPT on the B200 path (BASELINE config 4).

Mirrors the reference's `Lit_minGPT.training_step` / `shared_step` (transformer/minGPT.py:413-422: teacher-forced logits,
mean cross entropy), `configure_optimizers` (:618-665: AdamW, betas (0.9, 0.95), weight decay 0.01 on Linear weights, none on
biases / LayerNorm / embeddings / pos_emb) and the DDP gradient averaging of `GPT_VAE_train.py:172-174`.

The computation runs in libmgv (include/mgv.h, mgv_gpt_train_*): forward with the three dropouts, backward through the tcgen05
GEMMs (dgrad / wgrad), attention / LayerNorm / GELU / cross-entropy backward, fused AdamW.  There is no autograd graph:
`GPTTrainer.step` fills `p.grad` of every parameter directly.  The parameters of the module become views into ONE flat fp32
buffer (as do their gradients), whose layout is block-contiguous, so that data-parallel training all-reduces a few large
contiguous buckets over NCCL / NVLink while the backward of the next bucket is still running.
"""
import ctypes
import os

import torch

from .. import _lib


class FusedAdamW:
    """torch.optim.AdamW semantics on the flat parameter buffer of a GPTTrainer (one kernel per step, which also refreshes
    the bf16 weight copies the GEMMs read).  `param_groups` has the reference's two groups (decay 0.01 / 0.0)."""

    def __init__(self, trainer, lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.01):
        self.trainer = trainer
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        decay, no_decay = [], []
        for name, p in trainer.model.named_parameters():
            (decay if trainer.decay_flags[name] else no_decay).append(p)
        self.param_groups = [dict(params=decay, weight_decay=weight_decay, lr=lr, betas=betas, eps=eps),
                             dict(params=no_decay, weight_decay=0.0, lr=lr, betas=betas, eps=eps)]
        self.exp_avg = torch.zeros_like(trainer.flat_params)
        self.exp_avg_sq = torch.zeros_like(trainer.flat_params)
        self.step_count = 0

    def zero_grad(self, set_to_none=False):
        self.trainer.flat_grads.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        tr = self.trainer
        tr.wait_for_gradients()
        g = self.param_groups[0]
        self.step_count += 1
        with tr.model._on_device():
            _lib.check(_lib.load().mgv_gpt_train_adamw(
                tr.model._handle(), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), float(g["lr"]), float(g["betas"][0]),
                float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self.step_count, float(tr.grad_scale),
                _lib.stream_ptr(tr.flat_params.device)), "mgv_gpt_train_adamw")
        return loss

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, defaults=self.defaults)

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])


class GPTTrainer:
    """Owns the flat parameter / gradient buffers of a GPT / GPTClass module on one GPU and runs its training step.

        trainer = GPTTrainer(lit.transformer)            # after .to('cuda'); parameters become views of trainer.flat_params
        opt = trainer.configure_optimizers(lr)           # FusedAdamW with the reference's decay / no-decay groups
        loss = trainer.step(idx, cls, targets)           # forward + backward: p.grad filled (and all-reduced when distributed)
        opt.step()
    """

    def __init__(self, model, layers_per_bucket=None, process_group=None):
        self.model = model
        if layers_per_bucket is None:    # blocks per all-reduce bucket (MGV_TRAIN_BUCKET_LAYERS overrides the default)
            layers_per_bucket = int(os.environ.get("MGV_TRAIN_BUCKET_LAYERS", "4"))
        self.layers_per_bucket = max(1, int(layers_per_bucket))
        self.process_group = process_group
        self.grad_scale = 1.0
        self.allreduce = True            # False: keep the local gradients even when torch.distributed is initialised
        self.overlap = True              # False: one all-reduce after the whole backward (no overlap; for measurements)
        self._pending = []
        self.seed = 783435
        p0 = model.head.weight
        if not p0.is_cuda:
            raise RuntimeError("GPTTrainer: move the model to a B200 first (libmgv has no CPU path)")
        dev = p0.device
        L = _lib.load()
        with model._on_device():
            h = model._handle()                      # loads the current weights into the handle
            total = int(L.mgv_gpt_train_numel(h))
            self.flat_params = torch.zeros(total, dtype=torch.float32, device=dev)
            self.flat_grads = torch.zeros(total, dtype=torch.float32, device=dev)
            self.layout, self.decay_flags = {}, {}
            off, num, dec = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
            for name, p in model.named_parameters():
                _lib.check(L.mgv_gpt_train_layout(h, name.encode(), ctypes.byref(off), ctypes.byref(num), ctypes.byref(dec)),
                           "mgv_gpt_train_layout(%s)" % name)
                o, n = int(off.value), int(num.value)
                if n != p.numel():
                    raise RuntimeError("GPTTrainer: %s has %d elements, libmgv expects %d" % (name, p.numel(), n))
                self.layout[name] = (o, n)
                self.decay_flags[name] = bool(dec.value)
                view = self.flat_params[o:o + n].view(p.shape)
                view.copy_(p.data.float())
                p.data = view                                        # the parameter IS the flat slice from now on
                p.grad = self.flat_grads[o:o + n].view(p.shape)
            model._mgv_sig = None                                    # data pointers changed: reload once ...
            h = model._handle()
            _lib.check(L.mgv_gpt_train_bind(h, _lib.ptr(self.flat_params), _lib.ptr(self.flat_grads), _lib.stream_ptr(dev)),
                       "mgv_gpt_train_bind")
        self._bound_handle = model._mgv_handle

    # ---------------------------------------------------------------- reference :618-665
    def configure_optimizers(self, lr, betas=(0.9, 0.95), weight_decay=0.01):
        return FusedAdamW(self, lr=lr, betas=betas, weight_decay=weight_decay)

    def _buckets(self):
        """[(layer_hi, layer_lo, flat_lo, flat_hi)] from the last blocks to the first; the head rides with the first bucket,
        the embeddings with the last"""
        nl = self.model.config.n_layer
        total = self.flat_params.numel()
        out = []
        hi = nl
        while hi > 0:
            lo = max(0, hi - self.layers_per_bucket)
            f_lo = self.layout["blocks.%d.ln1.weight" % lo][0] if lo > 0 else 0
            f_hi = total if hi == nl else self.layout["blocks.%d.ln1.weight" % hi][0]
            out.append((hi, lo, f_lo, f_hi))
            hi = lo
        return out

    def wait_for_gradients(self):
        for w in self._pending:
            w.wait()
        self._pending = []

    @torch.no_grad()
    def step(self, idx, cls, targets):
        """idx (B, t) int64 tokens, cls (B, 1) / (B,) int64 class ids or None, targets (B, m + t) int64.
        Returns the mean cross-entropy loss (0-d device tensor); gradients are in p.grad / self.flat_grads."""
        import torch.distributed as dist
        m = self.model
        if m._mgv_handle is None or m._mgv_handle is not self._bound_handle:
            raise RuntimeError("GPTTrainer: the model's libmgv handle changed (moved to another device?); build a new GPTTrainer")
        dev = self.flat_params.device
        idx = idx.to(device=dev, dtype=torch.int64).contiguous()
        targets = targets.to(device=dev, dtype=torch.int64).contiguous()
        B, t = idx.shape
        mm = 0
        if cls is not None:
            cls = cls.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
            mm = 1
        if targets.numel() != B * (mm + t):
            raise RuntimeError("GPTTrainer.step: %d targets for %d rows" % (targets.numel(), B * (mm + t)))
        training = m.training
        p_embd = float(m.drop.p) if training else 0.0
        p_resid = float(m.blocks[0].attn.resid_drop.p) if training else 0.0
        p_attn = float(m.blocks[0].attn.attn_drop.p) if training else 0.0
        self.seed = (self.seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        self.last_seed = self.seed
        distributed = (self.allreduce and dist.is_available() and dist.is_initialized()
                       and dist.get_world_size(self.process_group) > 1)
        L = _lib.load()
        self.wait_for_gradients()
        with m._on_device():
            h = m._handle()
            st = _lib.stream_ptr(dev)
            loss = torch.zeros((), dtype=torch.float32, device=dev)
            _lib.check(L.mgv_gpt_train_forward(h, _lib.ptr(idx), B, t, _lib.ptr(cls), mm, _lib.ptr(targets), p_embd, p_resid,
                                               p_attn, self.seed, _lib.ptr(loss), st), "mgv_gpt_train_forward")
            for hi, lo, f_lo, f_hi in self._buckets():
                _lib.check(L.mgv_gpt_train_backward(h, hi, lo, st), "mgv_gpt_train_backward")
                if distributed and self.overlap:
                    # NCCL averages this bucket (DDP semantics) on its own stream while the next bucket's backward runs on ours
                    self._pending.append(dist.all_reduce(self.flat_grads[f_lo:f_hi], op=dist.ReduceOp.AVG,
                                                         group=self.process_group, async_op=True))
        if distributed and not self.overlap:
            self._pending.append(dist.all_reduce(self.flat_grads, op=dist.ReduceOp.AVG, group=self.process_group, async_op=True))
        if distributed:
            dist.all_reduce(loss, op=dist.ReduceOp.AVG, group=self.process_group)   # logging value (sync_dist)
        return loss
