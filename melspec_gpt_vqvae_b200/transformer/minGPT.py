"""Drop-in for the hot-path part of the reference's transformer/minGPT.py.

Same class names, constructor contracts (`args` namespace), method signatures and
state_dict keys as the reference (GPTConfig :30-40, CausalSelfAttention :45-90, Block
:93-119, GPT :121-199, GPTClass :203-212, Lit_minGPT :216-665) -- the computation runs in
libmgv (include/mgv.h):

  GPT.forward / GPTClass.forward   -> mgv_gpt_forward   (bf16 tcgen05 GEMMs, fp32 accumulate)
  Lit_minGPT.sample                -> mgv_gpt_generate  (KV cache + CUDA-graph decode loop)
  Lit_minGPT.decode_to_img         -> code_reader + mgv_vqvae_decode_codes
  GPT.forward(targets=...) loss    -> mgv_gpt_cross_entropy

The forward / sample paths are inference (eval mode) kernels; the teacher-forced training step lives in
transformer/train_step.py.  No CPU fallback.

Device handling: every libmgv call runs with the device that holds the parameters made current
(`torch.cuda.device(p.device)`), on that device's current torch stream, and the handle is rebuilt when the
parameters move to another GPU -- a model on cuda:1 works while cuda:0 is current, as with the reference.
"""
import ctypes
import logging

import numpy as np
import torch
import torch.nn as nn
from torch.nn import functional as F

from .. import _lib
from ..vqvae.big_model_attn_gan import LitVQVAE, _params_signature, _SigCacheMixin

try:
    import pytorch_lightning as pl
    _LitBase = pl.LightningModule
except Exception:  # pragma: no cover
    _LitBase = nn.Module

logger = logging.getLogger(__name__)


def _no_callback(k):
    """default `callback` of Lit_minGPT.sample (reference :293: `lambda k: None`)"""
    return None

_HOT_PATH_ONLY = ("this submodule only owns parameters; the computation runs fused inside libmgv -- call "
                  "GPT.forward / Lit_minGPT.sample instead")


class GPTConfig:
    """ base GPT config, params common to all GPT versions (reference :30-40) """
    embd_pdrop = 0.1
    resid_pdrop = 0.1
    attn_pdrop = 0.1

    def __init__(self, vocab_size, block_size, **kwargs):
        self.vocab_size = vocab_size
        self.block_size = block_size
        for k, v in kwargs.items():
            setattr(self, k, v)


class CausalSelfAttention(nn.Module):
    """Parameter container (reference :45-90): key / query / value / proj Linear layers and the
    persistent `mask` buffer."""

    def __init__(self, config):
        super().__init__()
        assert config.n_embd % config.n_head == 0
        self.key = nn.Linear(config.n_embd, config.n_embd)
        self.query = nn.Linear(config.n_embd, config.n_embd)
        self.value = nn.Linear(config.n_embd, config.n_embd)
        self.attn_drop = nn.Dropout(config.attn_pdrop)
        self.resid_drop = nn.Dropout(config.resid_pdrop)
        self.proj = nn.Linear(config.n_embd, config.n_embd)
        mask = torch.tril(torch.ones(config.block_size, config.block_size))
        if hasattr(config, "n_unmasked"):
            mask[:config.n_unmasked, :config.n_unmasked] = 1
        self.register_buffer("mask", mask.view(1, 1, config.block_size, config.block_size))
        self.n_head = config.n_head

    def forward(self, x, layer_past=None):
        raise NotImplementedError("CausalSelfAttention: " + _HOT_PATH_ONLY)


class Block(nn.Module):
    """Parameter container (reference :93-119)."""

    def __init__(self, config):
        super().__init__()
        self.ln1 = nn.LayerNorm(config.n_embd)
        self.ln2 = nn.LayerNorm(config.n_embd)
        self.attn = CausalSelfAttention(config)
        self.mlp = nn.Sequential(
            nn.Linear(config.n_embd, 4 * config.n_embd),
            nn.GELU(),
            nn.Linear(4 * config.n_embd, config.n_embd),
            nn.Dropout(config.resid_pdrop),
        )

    def forward(self, x):
        raise NotImplementedError("Block: " + _HOT_PATH_ONLY)


class GPT(_SigCacheMixin, nn.Module):
    """ the full GPT language model, with a context size of block_size (reference :121-199) """

    def __init__(self, args, embd_pdrop=0., resid_pdrop=0., attn_pdrop=0., n_unmasked=0, last_linear=None,
                 block_size=None):
        super().__init__()
        config = GPTConfig(vocab_size=args.vocab_size, block_size=args.block_size,
                           embd_pdrop=embd_pdrop, resid_pdrop=resid_pdrop, attn_pdrop=attn_pdrop,
                           n_layer=args.n_layer, n_head=args.n_head, n_embd=args.n_embd,
                           n_unmasked=n_unmasked, last_linear=last_linear)
        if block_size is not None:
            config.block_size = block_size
        self.tok_emb = nn.Embedding(config.vocab_size, config.n_embd)
        self.pos_emb = nn.Parameter(torch.zeros(1, config.block_size, config.n_embd))
        self.drop = nn.Dropout(config.embd_pdrop)
        self.blocks = nn.Sequential(*[Block(config) for _ in range(config.n_layer)])
        self.ln_f = nn.LayerNorm(config.n_embd)
        output_size = last_linear if config.last_linear is not None else config.vocab_size
        self.head = nn.Linear(config.n_embd, output_size, bias=False)
        self.block_size = config.block_size
        self.apply(self._init_weights)
        self.config = config
        self._mgv_handle = None
        self._mgv_sig = None
        self._mgv_dev = None
        self._deterministic = False
        self._sig_cache_reset()
        logger.info("number of parameters: %e", sum(p.numel() for p in self.parameters()))

    def get_block_size(self):
        return self.block_size

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(module, nn.Linear) and module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    # ---------------------------------------------------------------- libmgv handle
    def _class_size(self):
        emb = getattr(self, "embedder", None)
        return emb.num_embeddings if emb is not None else 0

    def _handle(self):
        """libmgv handle for the device that holds the parameters (call with that device current: `_on_device`)."""
        p = self.head.weight
        if not p.is_cuda:
            raise RuntimeError("GPT: parameters are on %s; libmgv has no CPU path -- call .to('cuda')" % p.device)
        L = _lib.load()
        if self._mgv_handle is not None and self._mgv_dev != p.device:
            self._release_handle()                     # the module moved to another GPU
        if self._mgv_handle is None:
            cfg = _lib.GptConfig(vocab_size=self.config.vocab_size, block_size=self.block_size,
                                 n_layer=self.config.n_layer, n_head=self.config.n_head, n_embd=self.config.n_embd,
                                 class_size=self._class_size(), n_unmasked=int(getattr(self.config, "n_unmasked", 0) or 0),
                                 head_out=0 if self.config.last_linear is None else int(self.head.out_features))
            h = ctypes.c_void_p()
            _lib.check(L.mgv_gpt_create(ctypes.byref(cfg), ctypes.byref(h)), "mgv_gpt_create")
            self._mgv_handle = h
            self._mgv_dev = p.device
            self._mgv_sig = None
            if self._deterministic:
                _lib.check(L.mgv_gpt_set_deterministic(h, 1), "mgv_gpt_set_deterministic")
        sig = _params_signature(self, self)
        if sig != self._mgv_sig:
            st = _lib.stream_ptr(p.device)
            for k, t in self.state_dict().items():
                if k.endswith("attn.mask"):
                    continue
                t32 = t.detach().to(torch.float32).contiguous()
                _lib.check(L.mgv_gpt_load_weight(self._mgv_handle, k.encode(), _lib.ptr(t32), t32.numel(), st),
                           "mgv_gpt_load_weight(%s)" % k)
            torch.cuda.current_stream(p.device).synchronize()
            self._mgv_sig = sig
        return self._mgv_handle

    def _on_device(self):
        """context manager: the parameters' GPU is the current device while libmgv is called"""
        p = self.head.weight
        if not p.is_cuda:
            raise RuntimeError("GPT: parameters are on %s; libmgv has no CPU path -- call .to('cuda')" % p.device)
        return torch.cuda.device(p.device)

    def refresh_weights(self):
        """Re-pack the bf16 weight copies inside libmgv on the next call.  Needed only after edits that bypass autograd's
        version counter (`p.data.copy_()`, `p.data.mul_()` ...): load_state_dict, .to() and ordinary in-place ops are
        detected automatically.  Also needed after re-binding an attribute to a new nn.Parameter object."""
        self._mgv_sig = None
        self._sig_cache_reset()

    @property
    def deterministic(self):
        return self._deterministic

    @deterministic.setter
    def deterministic(self, on):
        """True: decode without split-K reductions (bit-reproducible run to run like the reference's greedy path, slower)."""
        self._deterministic = bool(on)
        if self._mgv_handle is not None:
            with torch.cuda.device(self._mgv_dev):
                _lib.check(_lib.load().mgv_gpt_set_deterministic(self._mgv_handle, 1 if on else 0), "mgv_gpt_set_deterministic")

    def _release_handle(self):
        if getattr(self, "_mgv_handle", None) is not None:
            try:
                with torch.cuda.device(self._mgv_dev):
                    _lib.load().mgv_gpt_destroy(self._mgv_handle)
            finally:
                self._mgv_handle = None
                self._mgv_sig = None

    def __getstate__(self):
        # the libmgv handle is a process-local cache: drop it so that copy.deepcopy / pickle / torch.save(model) work
        state = self.__dict__.copy()
        state["_mgv_handle"] = None
        state["_mgv_sig"] = None
        state["_mgv_dev"] = None
        state["_mgv_sig_tensors"] = {}
        return state

    def __del__(self):
        try:
            self._release_handle()
        except Exception:
            pass

    def _check_inference(self):
        if self.training and (self.drop.p > 0 or any(b.attn.attn_drop.p > 0 or b.attn.resid_drop.p > 0 for b in self.blocks)):
            raise NotImplementedError("GPT.forward in training mode with dropout returns no autograd graph on the B200 path: "
                                      "use transformer/train_step.py (GPTTrainer.step / Lit_minGPT.training_step) to train, "
                                      "or call .eval() for inference")

    def _forward_impl(self, idx, embeddings, cls):
        self._check_inference()
        if not idx.is_cuda:
            raise RuntimeError("GPT.forward: idx is on %s; libmgv has no CPU path" % idx.device)
        idx = idx.to(torch.int64).contiguous()
        B, t = idx.shape
        m = 0
        emb = None
        if embeddings is not None:
            emb = embeddings.detach().to(torch.float32).contiguous()
            m = emb.shape[1]
        elif cls is not None:
            cls = cls.to(torch.int64).reshape(-1).contiguous()
            m = 1
        T = m + t
        assert T <= self.block_size, "Cannot forward, model block size is exhausted."      # reference :178
        vout = self.head.out_features
        nh = self.config.n_head
        dev = self.head.weight.device
        if idx.device != dev:
            raise RuntimeError("GPT.forward: idx is on %s but the parameters are on %s" % (idx.device, dev))
        with self._on_device():
            logits = torch.empty(B, T, vout, dtype=torch.float32, device=dev)
            att = torch.empty(B, nh, T, T, dtype=torch.float32, device=dev)
            _lib.check(_lib.load().mgv_gpt_forward(self._handle(), _lib.ptr(idx), B, t, _lib.ptr(emb), _lib.ptr(cls), m,
                                                   _lib.ptr(logits), _lib.ptr(att), _lib.stream_ptr(dev)), "mgv_gpt_forward")
        return logits, att

    @torch.no_grad()
    def forward(self, idx, embeddings=None, targets=None):
        """-> (logits (B, m+t, V), loss | None, att (B, n_head, m+t, m+t))   (reference :168-199)"""
        logits, att = self._forward_impl(idx, embeddings, None)
        loss = None
        if targets is not None:
            # F.cross_entropy (reference :197): mean over the targets that are not ignore_index (-100)
            t = targets.reshape(-1)
            keep = t != -100
            rows = self.cross_entropy_rows(logits.view(-1, logits.size(-1)), t.masked_fill(~keep, 0))
            loss = (rows * keep).sum() / keep.sum()
        return logits, loss, att

    def cross_entropy_rows(self, logits2d, targets1d):
        """per-row cross entropy (rows,) fp32 of fp32 logits (rows, V) -> mgv_gpt_cross_entropy"""
        if not logits2d.is_cuda:
            raise RuntimeError("cross_entropy_rows: logits are on %s; libmgv has no CPU path" % logits2d.device)
        lg = logits2d.detach().to(torch.float32).contiguous()
        t = targets1d.detach().to(device=lg.device, dtype=torch.int64).contiguous()
        rows, V = lg.shape
        assert t.numel() == rows, "cross_entropy_rows: %d targets for %d rows" % (t.numel(), rows)
        with self._on_device():
            out = torch.empty(rows, dtype=torch.float32, device=lg.device)
            _lib.check(_lib.load().mgv_gpt_cross_entropy(self._handle(), _lib.ptr(lg), _lib.ptr(t), rows, V, _lib.ptr(out),
                                                         _lib.stream_ptr(lg.device)), "mgv_gpt_cross_entropy")
        return out

    def last_launches(self):
        with self._on_device():
            return int(_lib.load().mgv_gpt_last_launches(self._handle()))


class GPTClass(GPT):
    """reference :203-212"""

    def __init__(self, args):
        super().__init__(args, embd_pdrop=args.embd_pdrop, resid_pdrop=args.resid_pdrop, attn_pdrop=args.attn_pdrop,
                         n_unmasked=args.n_unmasked, last_linear=args.last_linear, block_size=args.block_size)
        self.embedder = nn.Embedding(args.class_size, args.n_embd)

    @torch.no_grad()
    def forward(self, idx, token):
        if token.dim() != 2 or token.shape[1] != 1:
            # the reference embeds every class token and prepends all of them; only the single-token
            # conditioning it actually uses (get_c: (B,1)) is fused -- longer prefixes go through `embeddings`
            emb = F.embedding(token.to(self.embedder.weight.device), self.embedder.weight)
            return super().forward(idx, embeddings=emb)
        logits, att = self._forward_impl(idx, None, token)
        return logits, None, att


class Lit_minGPT(_LitBase):
    """reference :216-665.  `args` is the reference's merged argparse/config namespace."""

    def __init__(self, args, ckpt_path=None, ignore_keys=[], first_stage_key="image", cond_stage_key="depth",
                 downsample_cond_size=-1, pkeep=1.0):
        super().__init__()
        self.args = args
        self.transformer = GPTClass(args)
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)
        self.first_stage_key = first_stage_key
        self.cond_stage_key = cond_stage_key
        self.downsample_cond_size = downsample_cond_size
        self.pkeep = pkeep
        self.return_attention = True   # sample() returns the (B,nh,T,T) attention like the reference (:360)
        self.sample_seed = 783435      # Philox key of the device-side sampler; advanced after every sample() call
        self.record_step_logits = False   # sample() keeps the per-step logits in self.last_step_logits
        self.last_step_logits = None
        self.datamodule_loader()
        self.forward_shuffle_idx, self.backward_shuffle_idx = self.make_idx(5, 53)
        if getattr(self.args, "reconstruct_spec", "") != "":
            self.first_stage_model = LitVQVAE(num_embeddings=128, embedding_dim=256)          # reference :242
            self.first_stage_model.load_state_dict(torch.load(self.args.reconstruct_spec))
            self.first_stage_model.eval().to(self.args.device)

    def init_from_ckpt(self, path, ignore_keys=list()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        for k in list(sd.keys()):
            for ik in ignore_keys:
                if k.startswith(ik):
                    print("Deleting key {} from state_dict.".format(k))
                    del sd[k]
        self.load_state_dict(sd, strict=False)
        print(f"Restored from {path}")

    def datamodule_loader(self):
        """The dataset layer (datasets/datamodule.py) is outside the hot path; it is used when the
        reference's `datasets` package is importable, otherwise left unset."""
        self.data = None
        try:
            from datasets.datamodule import DataModule  # the reference's module, if on sys.path
        except Exception:
            return
        if not hasattr(self.args, "spec_dir_path"):
            return
        self.data = DataModule(batch_size=self.args.batch_size, spec_dir_path=self.args.spec_dir_path, mel_num=80,
                               spec_len=860, spec_crop_len=848, random_crop=False,
                               num_workers=getattr(self.args, "workers", 0))
        self.data.setup()

    # ---------------------------------------------------------------- hot path
    @torch.no_grad()
    def forward(self, x, c=None):
        """teacher-forced logits (reference :260-285): -> (logits (B, 265, V), target = x)"""
        z_indices = x
        target = z_indices
        logits, _, _ = self.transformer(z_indices[:, :-1], c)
        cond_size = c.size(-1)
        logits = logits[:, cond_size - 1:]
        return logits, target

    def top_k_logits(self, logits, k):
        """reference :287-291 (host-side helper; the sampler applies the same rule on the device)"""
        v, ix = torch.topk(logits, k)
        out = logits.clone()
        out[out < v[..., [-1]]] = -float('Inf')
        return out

    @torch.no_grad()
    def sample(self, x, c, steps, temperature=1.0, sample=False, top_k=None, callback=_no_callback):
        """reference :293-360.  Returns (x (B, t0+steps) int64, att (B, n_head, T, T) fp32 on CPU).

        The whole loop runs on the device (KV cache + CUDA graph) in one libmgv call.  A caller-supplied `callback` is
        invoked before every step exactly like the reference does (:332), which needs the host in the loop: the tokens
        are then generated one step per call (each call prefills the KV cache from the tokens so far, i.e. the
        reference's no-cache schedule) -- same tokens for the same seed, much slower; leave the default for throughput.
        `self.record_step_logits = True` additionally keeps the logits of every step in `self.last_step_logits`
        (steps, B, V) for parity checks."""
        block_size = self.transformer.get_block_size()
        assert not self.transformer.training
        if self.pkeep <= 0.0:
            raise NotImplementedError('Implement for GPTFeatsCLass')
        tr = self.transformer
        if not x.is_cuda:
            raise RuntimeError("sample: x is on %s; libmgv has no CPU path" % x.device)
        x = x.to(torch.int64).contiguous()
        B, t0 = x.shape
        if isinstance(tr, GPTClass):
            cond_size = c.size(-1)
            if cond_size != 1:
                raise NotImplementedError("sample: only single-token class conditioning (c of shape (B,1)) is supported")
            cls = c.to(device=x.device, dtype=torch.int64).reshape(-1).contiguous()
            m = 1
        else:
            cls, m = None, 0
        if steps == 0:
            raise UnboundLocalError("local variable 'att' referenced before assignment")   # as the reference (:360)
        seed = int(self.sample_seed) & 0xFFFFFFFFFFFFFFFF
        self.sample_seed = (int(self.sample_seed) * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        kw = dict(cls=cls, m=m, temperature=temperature, sample=sample, top_k=top_k, seed=seed)
        self.last_step_logits = None
        if callback is _no_callback:
            for k in range(steps):
                assert t0 + k + m <= block_size                                   # reference :336 / :342
            out, att, sl = self._generate(x, steps, self.return_attention, self.record_step_logits, **kw)
        else:
            out, att, logs = x, None, []
            for k in range(steps):
                callback(k)                                                       # reference :332
                assert out.shape[1] + m <= block_size                             # reference :336 / :342
                out, att, sl = self._generate(out, 1, self.return_attention and k == steps - 1, self.record_step_logits, **kw)
                logs.append(sl)
            sl = torch.cat(logs, 0) if self.record_step_logits else None
        self.last_step_logits = sl
        return out, (self._att_to_host(att) if att is not None else None)

    def _att_to_host(self, att):
        """`att.detach().cpu()` of the reference (:360) into PINNED host memory: the (B, n_head, T, T) map of a 64-clip batch
        is 288 MB, which a pageable copy moves at ~2 GB/s and a pinned one at PCIe rate.  The tensor comes from torch's
        caching host allocator (one cudaHostAlloc on the first call, recycled blocks afterwards) and is returned as it is:
        the caller owns it like the reference's CPU tensor, and there is no second 288 MB copy out of a staging buffer."""
        host = torch.empty(att.shape, dtype=att.dtype, pin_memory=True)
        host.copy_(att.detach(), non_blocking=True)
        torch.cuda.current_stream(att.device).synchronize()
        return host

    def _generate(self, x, steps, want_att, want_logits, cls, m, temperature, sample, top_k, seed):
        """one mgv_gpt_generate call: x (B,t0) -> (B,t0+steps); the Philox counter is (position, row), so a generation split
        over several calls with the same seed draws the same numbers as a single call"""
        tr = self.transformer
        B, t0 = x.shape
        nh = tr.config.n_head
        Tf = m + t0 + steps - 1
        dev = x.device
        with tr._on_device():
            out = torch.empty(B, t0 + steps, dtype=torch.int64, device=dev)
            att = torch.empty(B, nh, Tf, Tf, dtype=torch.float32, device=dev) if want_att else None
            sl = None
            L = _lib.load()
            h = tr._handle()
            if want_logits:
                sl = torch.empty(steps, B, tr.config.vocab_size, dtype=torch.float32, device=dev)
                _lib.check(L.mgv_gpt_set_step_logits(h, _lib.ptr(sl)), "mgv_gpt_set_step_logits")
            _lib.check(L.mgv_gpt_generate(
                h, _lib.ptr(x) if t0 > 0 else None, B, t0, None, _lib.ptr(cls), m, int(steps),
                float(temperature), 1 if sample else 0, int(top_k) if top_k is not None else 0, seed,
                _lib.ptr(out), _lib.ptr(att), 1, _lib.stream_ptr(dev)), "mgv_gpt_generate")
        return out, att, sl

    # ---------------------------------------------------------------- data plumbing (reference :387-411)
    def get_x(self, batch):
        x = batch['codes']
        x = x.permute(0, 2, 1)
        x = torch.flatten(x, start_dim=1)
        return x.to(self.args.device)

    def get_c(self, batch):
        c = batch["target"].unsqueeze(1)
        return c.to(self.args.device)

    def get_xc(self, batch, N=None):
        x = self.get_x(batch)
        c = self.get_c(batch)
        if N is not None:
            x = x[:N]
            c = c[:N]
        return x, c

    def shared_step(self, batch, batch_idx):
        x, c = self.get_xc(batch)
        logits, target = self(x, c)
        return self.transformer.cross_entropy_rows(logits.reshape(-1, logits.size(-1)), target.reshape(-1)).mean()

    # ---------------------------------------------------------------- training (reference :413-422, :618-665)
    def trainer(self):
        """the GPTTrainer that owns the flat parameter / gradient buffers (created on first use, after .to('cuda'))"""
        from .train_step import GPTTrainer
        if getattr(self, "_trainer", None) is None or self._trainer.model._mgv_handle is not self._trainer._bound_handle:
            self._trainer = GPTTrainer(self.transformer)
        return self._trainer

    def training_step(self, batch, batch_idx):
        """Teacher-forced cross entropy like the reference's shared_step -- but forward AND backward run here, in libmgv
        (there is no autograd graph): the returned loss is a detached device scalar and every `p.grad` is already filled
        (and all-reduced when torch.distributed is initialised).  Follow it with `optimizer.step()`; under Lightning use
        manual optimisation."""
        x, c = self.get_xc(batch)
        return self.trainer().step(x[:, :-1], c, x)

    def configure_optimizers(self):
        """AdamW with the reference's decay / no-decay parameter groups (:618-665), fused into one libmgv kernel"""
        return self.trainer().configure_optimizers(lr=self.args.learning_rate)

    def validation_step(self, batch, batch_idx):
        loss = self.shared_step(batch, batch_idx)
        if hasattr(self, "log") and _LitBase is not nn.Module:
            self.log("val/loss", loss, prog_bar=True, logger=True, on_step=True, on_epoch=True)
        return loss

    # ---------------------------------------------------------------- code order helpers (reference :431-456)
    def make_idx(self, H, W):
        idx = np.arange(H * W).reshape(H, W)
        idx = idx.T
        idx = torch.tensor(idx.ravel())
        return idx, torch.argsort(idx)

    def code_reader(self, x, reverse=False):
        B, L = x.shape
        L_idx = len(self.forward_shuffle_idx)
        if L != L_idx:
            raise NotImplementedError("code_reader: only %d-token clips are supported (got %d)" % (L_idx, L))
        idx = self.backward_shuffle_idx if reverse else self.forward_shuffle_idx
        return x[:, idx.to(x.device)]

    @torch.no_grad()
    def decode_to_img(self, index, zshape, stage='first'):
        """tokens (B,265) time-major -> mel (B,1,80,848)   (reference :515-528)"""
        if stage == 'first':
            index = self.code_reader(index, reverse=True)
        else:
            raise NotImplementedError
        bhwc = (zshape[0], zshape[2], zshape[3], zshape[1])
        if bhwc[1:] == (5, 53, 256) and hasattr(self.first_stage_model, "decode_codes"):
            return self.first_stage_model.decode_codes(index.reshape(bhwc[0], -1))   # gather fused into the decoder
        quant_z = self.first_stage_model._vq_vae.get_codebook_entry(index.reshape(-1), shape=bhwc)
        return self.first_stage_model.decode(quant_z)

    @torch.no_grad()
    def log_images(self, batch, temperature=None, top_k=None, callback=None, lr_interface=False, **kwargs):
        """reference :530-612: half-prompt continuation, unconditional sample, greedy sample, reconstruction."""
        log = dict()
        N = 1
        x, c = self.get_xc(batch, N)
        quant_z_shape = (x.shape[0], 256, 5, 53)
        cb = callback if callback is not None else (lambda k: None)
        half = x.shape[1] // 2
        t = temperature if temperature is not None else 1.0
        idx, att_half = self.sample(x[:, :half], c, steps=x.shape[1] - half, temperature=t, sample=True,
                                    top_k=top_k if top_k is not None else 100, callback=cb)
        log["samples_half"] = self.decode_to_img(idx, quant_z_shape)
        idx, att_nopix = self.sample(x[:, :0], c, steps=x.shape[1], temperature=t, sample=True,
                                     top_k=top_k if top_k is not None else 100, callback=cb)
        log["samples_nopix"] = self.decode_to_img(idx, quant_z_shape)
        idx, att_det = self.sample(x[:, :0], c, steps=x.shape[1], sample=False, callback=cb)
        log["samples_det"] = self.decode_to_img(idx, quant_z_shape)
        log["reconstructions"] = self.decode_to_img(x, quant_z_shape)
        log["att_half"], log["att_nopix"], log["att_det"] = att_half, att_nopix, att_det
        return log
