"""Drop-in for the reference's transformer/decoders.py (GPTDecoder :9-123): the GPT-VAE generator.

A causal GPT conditioned on the latent z, passed as a float prefix embedding (position 0).  The transformer, the
per-position cross entropy and the autoregressive sampler run in libmgv:

  forward / reconstruct_error  -> mgv_gpt_forward (prefix embeddings) + mgv_gpt_cross_entropy
  sample                       -> mgv_gpt_generate (KV cache, CUDA-graph decode loop, prefix embeddings)

Same constructor contract, method names and state_dict keys (`transformer.*`, `loss.weight`) as the reference.
"""
import torch
import torch.nn as nn

from .. import _lib
from .minGPT import GPT


class GPTDecoder(nn.Module):
    """GPT decoder with constant-length data and conditioning first element"""

    def __init__(self, args, embd_pdrop=0., resid_pdrop=0., attn_pdrop=0., n_unmasked=0, last_linear=None,
                 block_size=None):
        super().__init__()
        self.args = args
        self.transformer = GPT(args, embd_pdrop=embd_pdrop, resid_pdrop=resid_pdrop, attn_pdrop=attn_pdrop,
                               n_unmasked=n_unmasked, last_linear=last_linear, block_size=block_size)
        # unit class weights, reduction='none' (reference :20-21); kept as a module for the `loss.weight` buffer
        self.loss = nn.CrossEntropyLoss(weight=torch.ones(args.vocab_size), reduction='none')
        self.return_attention = True
        self.sample_seed = 783435

    @torch.no_grad()
    def forward(self, x, c=None):
        """x (B, T) int64, c (B, m, n_embd) float -> (logits (B, T, V) with row i = p(x_i | x_<i, c), target = x)"""
        logits, _, _ = self.transformer(x[:, :-1], c)
        cond_size = c.size(-2)
        return logits[:, cond_size - 1:], x

    @torch.no_grad()
    def reconstruct_error(self, x, z):
        """-sum_i log p(x_i | x_<i, z): x (B, T), z (B, n_sample, nz) -> (B, n_sample)   (reference :41-70)"""
        batch_size, seq_len = x.size()
        n_sample = z.size(1)
        logits, tgt = self(x, z)
        if n_sample == 1:
            tgt = tgt.contiguous().view(-1)
        else:
            tgt = tgt.unsqueeze(1).expand(batch_size, n_sample, seq_len).contiguous().view(-1)
        rows = logits.reshape(-1, logits.size(2))
        if not bool((self.loss.weight == 1).all()):
            raise NotImplementedError("GPTDecoder.loss: only the reference's unit class weights are supported")
        loss = self.transformer.cross_entropy_rows(rows, tgt)
        return loss.view(batch_size, n_sample, -1).sum(-1)

    def log_probability(self, x, z):
        return -self.reconstruct_error(x, z)

    def top_k_logits(self, logits, k):
        """host-side helper (reference :85-89); the device sampler applies the same rule"""
        v, ix = torch.topk(logits, k)
        out = logits.clone()
        out[out < v[..., [-1]]] = -float('Inf')
        return out

    @torch.no_grad()
    def sample(self, x, c, steps, temperature=1.0, sample=False, top_k=None, callback=lambda k: None):
        """reference :91-123: x (B, t0) int64 (t0 may be 0), c (B, m, n_embd) float prefix
        -> (x (B, t0+steps) int64, att (B, n_head, Tf, Tf) fp32 on the CPU)"""
        tr = self.transformer
        block_size = tr.get_block_size()
        assert not tr.training
        if not x.is_cuda:
            raise RuntimeError("sample: x is on %s; libmgv has no CPU path" % x.device)
        x = x.to(torch.int64).contiguous()
        emb = c.detach().to(device=x.device, dtype=torch.float32).contiguous()
        B, t0 = x.shape
        m = emb.size(-2)
        for k in range(steps):
            callback(k)
            assert t0 + k + m <= block_size            # reference :101-102
        if steps == 0:
            raise UnboundLocalError("local variable 'att' referenced before assignment")   # as the reference (:123)
        Tf = m + t0 + steps - 1
        seed = int(self.sample_seed) & 0xFFFFFFFFFFFFFFFF
        self.sample_seed = (int(self.sample_seed) * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        with tr._on_device():
            out = torch.empty(B, t0 + steps, dtype=torch.int64, device=x.device)
            att = torch.empty(B, tr.config.n_head, Tf, Tf, dtype=torch.float32, device=x.device) if self.return_attention else None
            _lib.check(_lib.load().mgv_gpt_generate(
                tr._handle(), _lib.ptr(x) if t0 > 0 else None, B, t0, _lib.ptr(emb), None, m, int(steps), float(temperature),
                1 if sample else 0, int(top_k) if top_k is not None else 0, seed, _lib.ptr(out), _lib.ptr(att), 1,
                _lib.stream_ptr(x.device)), "mgv_gpt_generate")
        return out, (att.detach().cpu() if att is not None else None)
