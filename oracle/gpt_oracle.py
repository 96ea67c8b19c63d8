"""ORACLE -- test infrastructure, not product code.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

CPU fp32 restatement (functional torch ops on a plain state_dict -- no nn.Module, no
KV cache) of the reference minGPT path of karchkha/MelSpec_GPT_VQVAE:
  CausalSelfAttention.forward  transformer/minGPT.py:72-90
  Block.forward                transformer/minGPT.py:107-119
  GPT.forward                  transformer/minGPT.py:168-199
  GPTClass.forward             transformer/minGPT.py:209-212
  Lit_minGPT.forward           transformer/minGPT.py:260-285
  Lit_minGPT.top_k_logits      transformer/minGPT.py:287-291
  Lit_minGPT.sample            transformer/minGPT.py:293-360
  Lit_minGPT.make_idx/code_reader  transformer/minGPT.py:431-456
The arithmetic itself lives in torch (the reference pins torch==1.13.1; this image has
2.11 -- operator semantics are unchanged).  Parity pin: tests/golden/gpt_*.npz hold outputs
of the UNMODIFIED reference modules on seeded synthetic weights/inputs
(tests/golden/make_golden.py); tests/test_oracle_cpu.py checks this file against them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


class GPTCfg:
    def __init__(self, vocab_size=128, block_size=266, n_layer=24, n_head=16, n_embd=1024, class_size=8,
                 n_unmasked=0, last_linear=None):
        self.vocab_size, self.block_size, self.n_layer = vocab_size, block_size, n_layer
        self.n_head, self.n_embd, self.class_size = n_head, n_embd, class_size
        self.n_unmasked, self.last_linear = n_unmasked, last_linear


def causal_mask(block_size, n_unmasked=0):
    mask = torch.tril(torch.ones(block_size, block_size))            # :65
    mask[:n_unmasked, :n_unmasked] = 1                                # :67-68
    return mask


def attention(sd, prefix, x, cfg, mask):
    B, T, C = x.shape
    nh = cfg.n_head
    k = F.linear(x, sd[prefix + "key.weight"], sd[prefix + "key.bias"]).view(B, T, nh, C // nh).transpose(1, 2)      # :76
    q = F.linear(x, sd[prefix + "query.weight"], sd[prefix + "query.bias"]).view(B, T, nh, C // nh).transpose(1, 2)  # :77
    v = F.linear(x, sd[prefix + "value.weight"], sd[prefix + "value.bias"]).view(B, T, nh, C // nh).transpose(1, 2)  # :78
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(k.size(-1)))                                                  # :81
    att = att.masked_fill(mask[:T, :T] == 0, float("-inf"))                                                          # :82
    att = F.softmax(att, dim=-1)                                                                                     # :83
    y = att @ v                                                                                                      # :84 (eval: no dropout)
    y = y.transpose(1, 2).contiguous().view(B, T, C)                                                                 # :85
    y = F.linear(y, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])                                            # :88
    return y, att


def block(sd, i, x, cfg, mask):
    p = "blocks.%d." % i
    res = x
    h = F.layer_norm(x, (cfg.n_embd,), sd[p + "ln1.weight"], sd[p + "ln1.bias"], 1e-5)   # :113
    h, att = attention(sd, p + "attn.", h, cfg, mask)                                     # :114
    x = res + h                                                                            # :115
    h = F.layer_norm(x, (cfg.n_embd,), sd[p + "ln2.weight"], sd[p + "ln2.bias"], 1e-5)
    h = F.linear(h, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])
    h = F.gelu(h)                                                                          # nn.GELU() = erf form (:102)
    h = F.linear(h, sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])
    x = x + h                                                                              # :117
    return x, att


@torch.no_grad()
def gpt_forward(sd, cfg, idx, embeddings=None, targets=None):
    """GPT.forward (:168-199), eval mode.  Returns (logits, loss, att)."""
    tok = F.embedding(idx, sd["tok_emb.weight"])                                           # :170
    if embeddings is not None:
        tok = torch.cat((embeddings, tok), dim=1)                                          # :175
    t = tok.shape[1]
    assert t <= cfg.block_size, "Cannot forward, model block size is exhausted."           # :178
    x = tok + sd["pos_emb"][:, :t, :]                                                      # :179-180
    mask = causal_mask(cfg.block_size, cfg.n_unmasked).to(tok.device)
    att = None
    for i in range(cfg.n_layer):
        x, att = block(sd, i, x, cfg, mask)                                                # :185
    x = F.layer_norm(x, (cfg.n_embd,), sd["ln_f.weight"], sd["ln_f.bias"], 1e-5)           # :186
    logits = F.linear(x, sd["head.weight"])                                                # :188
    loss = None
    if targets is not None:
        loss = F.cross_entropy(logits.view(-1, logits.size(-1)), targets.view(-1))         # :197
    return logits, loss, att


@torch.no_grad()
def gptclass_forward(sd, cfg, idx, token):
    """GPTClass.forward (:209-212)."""
    emb = F.embedding(token, sd["embedder.weight"])                                        # :210
    return gpt_forward(sd, cfg, idx, embeddings=emb)


@torch.no_grad()
def lit_forward(sd, cfg, x, c):
    """Lit_minGPT.forward (:260-285): teacher-forced logits and targets."""
    target = x
    logits, _, _ = gptclass_forward(sd, cfg, x[:, :-1], c)                                 # :279
    cond_size = c.size(-1)                                                                 # :281
    logits = logits[:, cond_size - 1:]                                                     # :283
    return logits, target


def top_k_logits(logits, k):
    """Lit_minGPT.top_k_logits (:287-291): ties with the k-th value survive."""
    v, _ = torch.topk(logits, k)
    out = logits.clone()
    out[out < v[..., [-1]]] = -float("Inf")
    return out


@torch.no_grad()
def sample(sd, cfg, x, c, steps, temperature=1.0, sample=False, top_k=None, callback=lambda k: None,
           generator=None, return_step_logits=False):
    """Lit_minGPT.sample (:293-360), GPTClass branch: full forward per step (no KV cache)."""
    block_size = cfg.block_size
    att = None
    step_logits = []
    for k in range(steps):
        callback(k)                                                                        # :332
        cond_size = c.size(-1)
        assert x.size(1) + cond_size <= block_size                                         # :336
        logits, _, att = gptclass_forward(sd, cfg, x, c)                                   # :340
        logits = logits[:, -1, :] / temperature                                            # :346
        if return_step_logits:
            step_logits.append(logits.clone())
        if top_k is not None:
            logits = top_k_logits(logits, top_k)                                           # :349
        probs = F.softmax(logits, dim=-1)                                                  # :351
        if sample:
            ix = torch.multinomial(probs, num_samples=1, generator=generator)              # :354
        else:
            _, ix = torch.topk(probs, k=1, dim=-1)                                         # :356
        x = torch.cat((x, ix), dim=1)                                                      # :358
    if return_step_logits:
        return x, att, torch.stack(step_logits, 1)
    return x, att


def make_idx(H, W):
    """Lit_minGPT.make_idx (:431-435)."""
    idx = np.arange(H * W).reshape(H, W).T.ravel()
    return idx, np.argsort(idx)


def code_reader(x, reverse=False, H=5, W=53):
    """Lit_minGPT.code_reader (:438-456) for L == H*W."""
    fwd, bwd = make_idx(H, W)
    return x[:, bwd] if reverse else x[:, fwd]
