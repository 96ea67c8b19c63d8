"""GPU parity of the minGPT training step (BASELINE config 4): loss and every parameter gradient against the reference
math under torch autograd (fp32, CPU), with dropout off and with dropout on (the kernels' own counter-hash masks are
rebuilt and fed to the reference math), the fused AdamW against torch.optim.AdamW, and a full-size smoke run.
reference: Lit_minGPT.training_step / shared_step transformer/minGPT.py:413-422, configure_optimizers :618-665.
Tolerances: bf16 GEMM operands with fp32 accumulation -> per-tensor relative L2 error of a gradient <= 4e-2."""
import argparse
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from make_golden import GPT_SMALL
from melspec_gpt_vqvae_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu

GRAD_REL_TOL = 4e-2


def _lit(cfg, sd, pdrop=0.0):
    from melspec_gpt_vqvae_b200.transformer.minGPT import Lit_minGPT
    args = argparse.Namespace(embd_pdrop=pdrop, resid_pdrop=pdrop, attn_pdrop=pdrop, reconstruct_spec="", device="cuda",
                              learning_rate=1e-3, **cfg)
    lit = Lit_minGPT(args)
    missing = lit.transformer.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    return lit.to("cuda")


def _mask(seed, stream, p, shape):
    n = int(np.prod(shape))
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().mgv_test_dropout_mask(seed, stream, p, n, _lib.ptr(out), _lib.stream_ptr()), "mgv_test_dropout_mask")
    thresh = min(int(p * 16777216.0 + 0.5), 16777215)
    inv_keep = 16777216.0 / (16777216 - thresh)
    return out.view(shape).cpu().double() * inv_keep


def reference_loss(params, cfg, idx, cls, targets, masks=None):
    """GPTClass.forward + F.cross_entropy in double precision with explicit dropout masks (reference :72-90, :107-119,
    :168-199, :209-212, :416); `params` = dict of leaf tensors with requires_grad"""
    C, nh, L = cfg["n_embd"], cfg["n_head"], cfg["n_layer"]
    tok = F.embedding(idx, params["tok_emb.weight"])
    tok = torch.cat((F.embedding(cls, params["embedder.weight"]), tok), dim=1)
    B, T, _ = tok.shape
    x = tok + params["pos_emb"][:, :T, :]
    if masks:
        x = x * masks["embd"].view(B, T, C)
    causal = torch.tril(torch.ones(T, T, dtype=torch.bool))
    for l in range(L):
        p = "blocks.%d." % l
        h = F.layer_norm(x, (C,), params[p + "ln1.weight"], params[p + "ln1.bias"], 1e-5)
        q = F.linear(h, params[p + "attn.query.weight"], params[p + "attn.query.bias"]).view(B, T, nh, C // nh).transpose(1, 2)
        k = F.linear(h, params[p + "attn.key.weight"], params[p + "attn.key.bias"]).view(B, T, nh, C // nh).transpose(1, 2)
        v = F.linear(h, params[p + "attn.value.weight"], params[p + "attn.value.bias"]).view(B, T, nh, C // nh).transpose(1, 2)
        att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(C // nh))
        att = F.softmax(att.masked_fill(~causal, float("-inf")), dim=-1)
        if masks:
            att = att * masks["attn%d" % l].view(B, nh, T, T)
        y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
        y = F.linear(y, params[p + "attn.proj.weight"], params[p + "attn.proj.bias"])
        if masks:
            y = y * masks["resid_attn%d" % l].view(B, T, C)
        x = x + y
        h = F.layer_norm(x, (C,), params[p + "ln2.weight"], params[p + "ln2.bias"], 1e-5)
        h = F.gelu(F.linear(h, params[p + "mlp.0.weight"], params[p + "mlp.0.bias"]))
        h = F.linear(h, params[p + "mlp.2.weight"], params[p + "mlp.2.bias"])
        if masks:
            h = h * masks["resid_mlp%d" % l].view(B, T, C)
        x = x + h
    x = F.layer_norm(x, (C,), params["ln_f.weight"], params["ln_f.bias"], 1e-5)
    logits = F.linear(x, params["head.weight"])
    return F.cross_entropy(logits.view(-1, logits.size(-1)), targets.view(-1))


def _compare_grads(trainer, ref_params, what):
    worst = ("", 0.0)
    for name, p in trainer.model.named_parameters():
        g = p.grad.detach().cpu().double()
        r = ref_params[name].grad
        if float(r.abs().max()) < 1e-9:
            # analytically zero gradient (the key bias shifts every score of a row equally: softmax is invariant to it)
            assert float(g.abs().max()) < 2e-4, "%s: gradient of %s should vanish, max |g| = %.3e" % (what, name, float(g.abs().max()))
            continue
        rel = float((g - r).norm() / r.norm())
        if rel > worst[1]:
            worst = (name, rel)
        assert rel <= GRAD_REL_TOL, "%s: gradient of %s differs: rel L2 %.4f (|ref| %.3e)" % (what, name, rel, float(r.norm()))
    print("%s: worst gradient rel L2 error %.4f (%s)" % (what, worst[1], worst[0]))


@pytest.mark.parametrize("pdrop", [0.0, 0.3])
def test_loss_and_gradients_vs_autograd(pdrop):
    cfg = GPT_SMALL
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=101, perturb=True)
    lit = _lit(cfg, sd, pdrop).train()
    B, T = 3, 41                                    # R = 123 rows: not a multiple of 64 (padded contraction)
    g = torch.Generator().manual_seed(5)
    x = torch.randint(0, 128, (B, T), generator=g)
    c = torch.randint(0, 8, (B, 1), generator=g)
    tr = lit.trainer()
    loss = tr.step(x[:, :-1].cuda(), c.cuda(), x.cuda())
    masks = None
    if pdrop > 0:
        seed, C, nh, R = tr.last_seed, cfg["n_embd"], cfg["n_head"], B * T
        masks = {"embd": _mask(seed, 1, pdrop, (R, C))}
        for l in range(cfg["n_layer"]):
            masks["attn%d" % l] = _mask(seed, 16 + 4 * l, pdrop, (B * nh, T, T))
            masks["resid_attn%d" % l] = _mask(seed, 17 + 4 * l, pdrop, (R, C))
            masks["resid_mlp%d" % l] = _mask(seed, 18 + 4 * l, pdrop, (R, C))
        keep = float((masks["embd"] > 0).double().mean())
        assert abs(keep - (1 - pdrop)) < 0.02, "dropout keep rate %.3f" % keep
    with torch.enable_grad():       # (importing make_golden switches autograd off globally)
        ref_params = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
        ref = reference_loss(ref_params, cfg, x[:, :-1], c, x, masks)
        ref.backward()
    print("pdrop %.1f: loss %.6f vs reference %.6f" % (pdrop, float(loss), float(ref)))
    assert abs(float(loss) - float(ref)) <= 5e-3 * max(1.0, abs(float(ref)))
    _compare_grads(tr, ref_params, "pdrop=%.1f" % pdrop)


def test_fused_adamw_matches_torch():
    cfg = GPT_SMALL
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=101, perturb=True)
    lit = _lit(cfg, sd).train()
    tr = lit.trainer()
    opt = lit.configure_optimizers()
    assert [g["weight_decay"] for g in opt.param_groups] == [0.01, 0.0]
    n_decay = sum(p.numel() for p in opt.param_groups[0]["params"])
    C, L = cfg["n_embd"], cfg["n_layer"]
    assert n_decay == L * 12 * C * C + cfg["vocab_size"] * C        # Linear weights only (reference :629-647)
    ref_p = {k: v.detach().clone().cpu() for k, v in lit.transformer.named_parameters()}
    for v in ref_p.values():
        v.requires_grad_(True)
    decay = [ref_p[k] for k in ref_p if tr.decay_flags[k]]
    nodecay = [ref_p[k] for k in ref_p if not tr.decay_flags[k]]
    topt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.01}, {"params": nodecay, "weight_decay": 0.0}], lr=1e-3,
                             betas=(0.9, 0.95))
    g = torch.Generator().manual_seed(6)
    for it in range(3):
        x = torch.randint(0, 128, (2, 30), generator=g)
        c = torch.randint(0, 8, (2, 1), generator=g)
        tr.step(x[:, :-1].cuda(), c.cuda(), x.cuda())
        for k, p in lit.transformer.named_parameters():
            ref_p[k].grad = p.grad.detach().cpu().clone()
        opt.step()
        topt.step()
        for k, p in lit.transformer.named_parameters():
            d = float((p.detach().cpu() - ref_p[k].detach()).abs().max())
            assert d <= 2e-6, "AdamW step %d: %s differs by %.3e" % (it, k, d)
    # the inference copies were refreshed by the optimizer: the eval forward sees the updated weights
    lit.eval()
    xq = torch.randint(0, 128, (2, 12), generator=g)
    cq = torch.randint(0, 8, (2, 1), generator=g)
    a, _ = lit(xq.cuda(), cq.cuda())
    lit.transformer.refresh_weights()
    b, _ = lit(xq.cuda(), cq.cuda())
    assert float((a - b).abs().max()) < 1e-5


def test_training_reduces_loss_and_generation_follows():
    cfg = GPT_SMALL
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=101, perturb=False)
    lit = _lit(cfg, sd, pdrop=0.1).train()
    lit.args.learning_rate = 3e-3
    opt = lit.configure_optimizers()
    g = torch.Generator().manual_seed(7)
    codes = torch.randint(0, 128, (4, 5, 53), generator=g)
    batch = {"codes": codes, "target": torch.tensor([1, 2, 3, 4])}
    losses = []
    for it in range(60):
        losses.append(float(lit.training_step(batch, it)))
        opt.step()
    print("loss %.3f -> %.3f over 60 steps on one batch" % (losses[0], losses[-1]))
    assert losses[0] > 4.5 and losses[-1] < 0.5 * losses[0]
    # the memorised clips come back from greedy generation (weights flow from the optimizer into the decode chain)
    lit.eval()
    lit.return_attention = False
    x = lit.get_x(batch)
    xs, _ = lit.sample(torch.zeros(4, 0, dtype=torch.long, device="cuda"), lit.get_c(batch), steps=265)
    assert float((xs == x).float().mean()) > 0.9


def test_config4_size_step_runs():
    """VGGSound-derived config (class_size 309), per-GPU batch 8 (config_GPT_vas.py:10), dropout 0.5: shapes, finiteness,
    loss at initialisation ~ ln(128), time per step."""
    cfg = dict(synthetic.GPT_VAS, class_size=309)
    sd = synthetic.synthetic_gpt_state_dict(cfg, seed=783435, perturb=False)
    lit = _lit(cfg, sd, pdrop=0.5).train()
    lit.args.learning_rate = 1e-6
    opt = lit.configure_optimizers()
    g = torch.Generator().manual_seed(8)
    batch = {"codes": torch.randint(0, 128, (8, 5, 53), generator=g), "target": torch.randint(0, 309, (8,), generator=g)}
    for it in range(3):
        loss = lit.training_step(batch, it)
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5):
        loss = lit.training_step(batch, it)
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("config 4 (24L/1024, bs=8 x 265, dropout 0.5): %.2f ms per training step, loss %.4f" % (ms, float(loss)))
    assert math.isfinite(float(loss)) and abs(float(loss) - math.log(128)) < 0.5
    assert all(torch.isfinite(p.grad).all() for p in lit.transformer.parameters())
