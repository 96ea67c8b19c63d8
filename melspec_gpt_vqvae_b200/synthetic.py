"""Deterministic synthetic weights / inputs for tests and bench (no checkpoints ship with the
reference and there is no network).  The key names and shapes are the reference's
state_dict layout (SURVEY.md section 5 "Checkpoint"): they load into the reference modules
and into the drop-ins alike.

Distributions follow the reference's initialisers (minGPT.py:159-166 N(0, 0.02) /
LayerNorm (1, 0); big_model_attn_gan.py:16 codebook U(+-1/K); torch conv default
U(+-1/sqrt(fan_in))).  With `perturb=True` the normally-constant tensors (LayerNorm /
GroupNorm affine, biases, pos_emb) get small random values so tests exercise them.
"""
import math
from collections import OrderedDict

import torch

GPT_VAS = dict(vocab_size=128, block_size=266, n_layer=24, n_head=16, n_embd=1024, class_size=8,
               n_unmasked=0, last_linear=None)  # config/config_GPT_vas.py

VQVAE_CH = 128
VQVAE_CH_MULT = [1, 1, 2, 2, 4]
VQVAE_NUM_RES_BLOCKS = 2
VQVAE_Z = 256


def gpt_param_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    """state_dict keys of GPTClass (or GPT when class_size == 0) in registration order, minus the
    `attn.mask` buffers."""
    C, V = cfg["n_embd"], cfg["vocab_size"]
    out_size = cfg.get("last_linear") or V
    s = OrderedDict()
    s["pos_emb"] = (1, cfg["block_size"], C)
    s["tok_emb.weight"] = (V, C)
    for i in range(cfg["n_layer"]):
        p = "blocks.%d." % i
        s[p + "ln1.weight"] = (C,)
        s[p + "ln1.bias"] = (C,)
        s[p + "ln2.weight"] = (C,)
        s[p + "ln2.bias"] = (C,)
        for n in ("key", "query", "value", "proj"):
            s[p + "attn.%s.weight" % n] = (C, C)
            s[p + "attn.%s.bias" % n] = (C,)
        s[p + "mlp.0.weight"] = (4 * C, C)
        s[p + "mlp.0.bias"] = (4 * C,)
        s[p + "mlp.2.weight"] = (C, 4 * C)
        s[p + "mlp.2.bias"] = (C,)
    s["ln_f.weight"] = (C,)
    s["ln_f.bias"] = (C,)
    s["head.weight"] = (out_size, C)
    if cfg.get("class_size", 0):
        s["embedder.weight"] = (cfg["class_size"], C)
    return s


def synthetic_gpt_state_dict(cfg: dict, seed: int = 783435, perturb: bool = True, with_mask: bool = False):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in gpt_param_shapes(cfg).items():
        if name.endswith("ln1.weight") or name.endswith("ln2.weight") or name == "ln_f.weight":
            t = torch.ones(shape)
            if perturb:
                t += 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias") or name == "pos_emb":
            t = torch.zeros(shape)
            if perturb:
                t += 0.02 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[name] = t
    if with_mask:
        bs, nu = cfg["block_size"], cfg.get("n_unmasked", 0)
        mask = torch.tril(torch.ones(bs, bs))
        mask[:nu, :nu] = 1
        for i in range(cfg["n_layer"]):
            sd["blocks.%d.attn.mask" % i] = mask.view(1, 1, bs, bs).clone()
    return sd


# ---------------------------------------------------------------------------------- VQVAE
def _resblock(s, p, cin, cout):
    s[p + ".norm1.weight"] = (cin,)
    s[p + ".norm1.bias"] = (cin,)
    s[p + ".conv1.weight"] = (cout, cin, 3, 3)
    s[p + ".conv1.bias"] = (cout,)
    s[p + ".norm2.weight"] = (cout,)
    s[p + ".norm2.bias"] = (cout,)
    s[p + ".conv2.weight"] = (cout, cout, 3, 3)
    s[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        s[p + ".nin_shortcut.weight"] = (cout, cin, 1, 1)
        s[p + ".nin_shortcut.bias"] = (cout,)


def _attnblock(s, p, c):
    s[p + ".norm.weight"] = (c,)
    s[p + ".norm.bias"] = (c,)
    for n in ("q", "k", "v", "proj_out"):
        s[p + ".%s.weight" % n] = (c, c, 1, 1)
        s[p + ".%s.bias" % n] = (c,)


def vqvae_param_shapes(num_embeddings=128, embedding_dim=256, encoder=True, decoder=True):
    """state_dict keys of LitVQVAE (big_model_attn_gan.py:538-602) except `discriminator.*`."""
    ch, mult, nrb = VQVAE_CH, VQVAE_CH_MULT, VQVAE_NUM_RES_BLOCKS
    nres = len(mult)
    s = OrderedDict()
    if encoder:
        e = "_encoder"
        s[e + ".conv_in.weight"] = (ch, 1, 3, 3)
        s[e + ".conv_in.bias"] = (ch,)
        in_mult = [1] + mult
        block_in = ch
        for lvl in range(nres):
            block_in = ch * in_mult[lvl]
            block_out = ch * mult[lvl]
            for b in range(nrb):
                _resblock(s, "%s.down.%d.block.%d" % (e, lvl, b), block_in, block_out)
                block_in = block_out
            if lvl == nres - 1:   # module order: down.block.* precede down.attn.*
                for b in range(nrb):
                    _attnblock(s, "%s.down.%d.attn.%d" % (e, lvl, b), block_in)
            if lvl != nres - 1:
                s["%s.down.%d.downsample.conv.weight" % (e, lvl)] = (block_in, block_in, 3, 3)
                s["%s.down.%d.downsample.conv.bias" % (e, lvl)] = (block_in,)
        _resblock(s, e + ".mid.block_1", block_in, block_in)
        _attnblock(s, e + ".mid.attn_1", block_in)
        _resblock(s, e + ".mid.block_2", block_in, block_in)
        s[e + ".norm_out.weight"] = (block_in,)
        s[e + ".norm_out.bias"] = (block_in,)
        s[e + ".conv_out.weight"] = (VQVAE_Z, block_in, 3, 3)
        s[e + ".conv_out.bias"] = (VQVAE_Z,)
    s["_vq_vae._embedding.weight"] = (num_embeddings, embedding_dim)
    if decoder:
        d = "_decoder"
        block_in = ch * mult[-1]
        s[d + ".conv_in.weight"] = (block_in, VQVAE_Z, 3, 3)
        s[d + ".conv_in.bias"] = (block_in,)
        _resblock(s, d + ".mid.block_1", block_in, block_in)
        _attnblock(s, d + ".mid.attn_1", block_in)
        _resblock(s, d + ".mid.block_2", block_in, block_in)
        ups = {}
        for lvl in reversed(range(nres)):
            block_out = ch * mult[lvl]
            keys = OrderedDict()
            for b in range(nrb + 1):
                _resblock(keys, "%s.up.%d.block.%d" % (d, lvl, b), block_in, block_out)
                block_in = block_out
            if lvl == nres - 1:
                for b in range(nrb + 1):
                    _attnblock(keys, "%s.up.%d.attn.%d" % (d, lvl, b), block_in)
            if lvl != 0:
                keys["%s.up.%d.upsample.conv.weight" % (d, lvl)] = (block_in, block_in, 3, 3)
                keys["%s.up.%d.upsample.conv.bias" % (d, lvl)] = (block_in,)
            ups[lvl] = keys
        for lvl in range(nres):  # ModuleList order: up.0 ... up.4
            s.update(ups[lvl])
        s[d + ".norm_out.weight"] = (block_in,)
        s[d + ".norm_out.bias"] = (block_in,)
        s[d + ".conv_out.weight"] = (1, block_in, 3, 3)
        s[d + ".conv_out.bias"] = (1,)
    if encoder:
        s["quant_conv.weight"] = (embedding_dim, VQVAE_Z, 1, 1)
        s["quant_conv.bias"] = (embedding_dim,)
    if decoder:
        s["post_quant_conv.weight"] = (VQVAE_Z, embedding_dim, 1, 1)
        s["post_quant_conv.bias"] = (VQVAE_Z,)
    return s


def synthetic_vqvae_state_dict(num_embeddings=128, embedding_dim=256, seed: int = 783435, perturb: bool = True,
                               encoder=True, decoder=True, codebook_scale=None):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in vqvae_param_shapes(num_embeddings, embedding_dim, encoder, decoder).items():
        if name == "_vq_vae._embedding.weight":
            if codebook_scale is None:
                t = (torch.rand(shape, generator=g) * 2 - 1) / num_embeddings       # U(-1/K, 1/K)
            else:
                t = torch.randn(shape, generator=g) * codebook_scale                # "trained-scale" codebook
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        elif ".norm" in name and name.endswith(".weight"):
            t = torch.ones(shape)
            if perturb:
                t += 0.1 * torch.randn(shape, generator=g)
        elif ".norm" in name and name.endswith(".bias"):
            t = torch.zeros(shape)
            if perturb:
                t += 0.05 * torch.randn(shape, generator=g)
        else:  # conv biases
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        sd[name] = t
    return sd


def melgan_param_shapes(n_mel=80, ngf=32, n_residual_layers=3):
    """state_dict keys / shapes of the reference's MelGAN Generator (vocoder/modules.py:38-77)."""
    shapes = OrderedDict()

    def conv(prefix, cout, cin, k, transposed=False):
        first = cin if transposed else cout
        shapes[prefix + ".bias"] = (cout,)
        shapes[prefix + ".weight_g"] = (first, 1, 1)
        shapes[prefix + ".weight_v"] = (cin, cout, k) if transposed else (cout, cin, k)

    mult = 16
    conv("model.1", mult * ngf, n_mel, 7)
    i = 2
    for r in (8, 8, 2, 2):
        conv("model.%d" % (i + 1), mult * ngf // 2, mult * ngf, 2 * r, transposed=True)
        for j in range(n_residual_layers):
            p = "model.%d" % (i + 2 + j)
            conv(p + ".block.2", mult * ngf // 2, mult * ngf // 2, 3)
            conv(p + ".block.4", mult * ngf // 2, mult * ngf // 2, 1)
            conv(p + ".shortcut", mult * ngf // 2, mult * ngf // 2, 1)
        i += 2 + n_residual_layers
        mult //= 2
    conv("model.%d" % (i + 2), 1, ngf, 7)
    return shapes


def synthetic_melgan_state_dict(n_mel=80, ngf=32, n_residual_layers=3, seed: int = 783435):
    """Random MelGAN weights at a trained-like scale: v ~ U(+-1/sqrt(fan_in)), g = |v| * U(1.4, 2.2) (the effective weight
    is g v / |v|, so g sets each filter's gain), small biases."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    shapes = melgan_param_shapes(n_mel, ngf, n_residual_layers)
    for name, shape in shapes.items():
        if name.endswith("weight_v"):
            fan_in = shape[1] * shape[2]
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
    for name, shape in shapes.items():
        if name.endswith("weight_g"):
            v = sd[name[:-1] + "v"]
            sd[name] = v.flatten(1).norm(dim=1).reshape(shape) * (1.4 + 0.8 * torch.rand(shape, generator=g))
        elif name.endswith("bias"):
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return OrderedDict((k, sd[k]) for k in shapes)
