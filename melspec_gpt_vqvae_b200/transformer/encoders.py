"""Drop-in for the reference's transformer/encoders.py (GPTEncoder :11-165): the GPT-VAE inference network.

A fully unmasked GPT (n_unmasked = block_size) whose head emits 2 * n_embd values; the last position's output is
split into (mean, logvar) of q(z | x).  The transformer runs in libmgv (mgv_gpt_forward, prefix-unmasked attention,
`last_linear` head); the Gaussian bookkeeping below is a handful of (B, nz) element-wise torch ops, as in the
reference.  Same constructor contract, method names, return values and state_dict keys (`transformer.*`).
"""
import math

import torch
import torch.nn as nn

from .minGPT import GPT
from .utils import log_sum_exp


class GPTEncoder(nn.Module):
    """GPT encoder with constant-length data"""

    def __init__(self, args, embd_pdrop=0., resid_pdrop=0., attn_pdrop=0., n_unmasked=0, last_linear=None,
                 block_size=None):
        super().__init__()
        self.args = args
        self.transformer = GPT(args, embd_pdrop=embd_pdrop, resid_pdrop=resid_pdrop, attn_pdrop=attn_pdrop,
                               n_unmasked=n_unmasked, last_linear=last_linear, block_size=block_size)

    @torch.no_grad()
    def forward(self, input):
        """input (B, T) int64 -> (mean (B, nz), logvar (B, nz), att (B, n_head, T, T))   (reference :21-45)"""
        logits, _, att = self.transformer.forward(input)
        mean, logvar = logits[:, -1, :].chunk(2, -1)
        if self.args.fix_var > 0:      # variance pinned to a constant (:41-43)
            logvar = mean.new_tensor([[[math.log(self.args.fix_var)]]]).expand_as(mean)
        return mean, logvar, att

    def encode_stats(self, x):
        return self.forward(x)

    def sample(self, input, nsamples):
        """-> z (B, nsamples, nz), (mu, logvar), att   (reference :51-63)"""
        mu, logvar, att = self.forward(input)
        return self.reparameterize(mu, logvar, nsamples), (mu, logvar), att

    def encode(self, input, nsamples):
        """-> z (B, nsamples, nz), KL(q(z|x) || N(0, I)) (B,)   (reference :65-83)"""
        mu, logvar, _ = self.forward(input)
        z = self.reparameterize(mu, logvar, nsamples)
        KL = 0.5 * (mu.pow(2) + logvar.exp() - logvar - 1).sum(dim=1)
        return z, KL

    def reparameterize(self, mu, logvar, nsamples=1):
        """mu + eps * exp(logvar / 2) with eps ~ N(0, I): (B, nz) -> (B, nsamples, nz)   (reference :85-108)"""
        B, nz = mu.size()
        std = logvar.mul(0.5).exp()
        eps = torch.zeros(B, nsamples, nz, dtype=std.dtype, device=std.device).normal_()
        return mu.unsqueeze(1) + eps * std.unsqueeze(1)

    def eval_inference_dist(self, x, z, param=None):
        """log q(z | x) for z (B, nsamples, nz) -> (B, nsamples)   (reference :110-138)"""
        nz = z.size(2)
        if not param:
            mu, logvar, _ = self.forward(x)
        else:
            mu, logvar = param
        mu, logvar = mu.unsqueeze(1), logvar.unsqueeze(1)
        dev = z - mu
        return -0.5 * ((dev ** 2) / logvar.exp()).sum(dim=-1) - 0.5 * (nz * math.log(2 * math.pi) + logvar.sum(-1))

    def calc_mi(self, x):
        """I(x, z) = E_x E_q(z|x) log q(z|x) - E_x E_q(z|x) log q(z), aggregate posterior over the batch (:140-165)"""
        mu, logvar, _ = self.forward(x)
        x_batch, nz = mu.size()
        neg_entropy = (-0.5 * nz * math.log(2 * math.pi) - 0.5 * (1 + logvar).sum(-1)).mean()
        z_samples = self.reparameterize(mu, logvar, 1)                  # (z_batch, 1, nz)
        mu, logvar = mu.unsqueeze(0), logvar.unsqueeze(0)               # (1, x_batch, nz)
        dev = z_samples - mu                                            # (z_batch, x_batch, nz)
        log_density = -0.5 * ((dev ** 2) / logvar.exp()).sum(dim=-1) - \
            0.5 * (nz * math.log(2 * math.pi) + logvar.sum(-1))
        log_qz = log_sum_exp(log_density, dim=1) - math.log(x_batch)
        return (neg_entropy - log_qz.mean(-1)).item()
