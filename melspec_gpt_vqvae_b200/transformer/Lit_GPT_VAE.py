"""Drop-in for the inference part of the reference's transformer/Lit_GPT_VAE.py (GPT_VAE :23-260): a VAE whose
inference network and generator are GPTs (GPTEncoder / GPTDecoder).  encode / decode / reconstruct / loss keep the
reference's signatures; the Lightning training / logging hooks and the optimiser set-up are not on the hot path
 and are not provided.
"""
import numpy as np
import torch
import torch.nn as nn

from .decoders import GPTDecoder
from .encoders import GPTEncoder

try:
    import pytorch_lightning as pl
    _LitBase = pl.LightningModule
except Exception:  # pragma: no cover
    _LitBase = nn.Module


class GPT_VAE(_LitBase):
    """VAE with normal prior"""

    def __init__(self, args):
        super().__init__()
        self.args = args
        # reference :42-43: unmasked encoder with a 2*n_embd head; decoder with one extra position for z
        self.encoder = GPTEncoder(args, n_unmasked=args.block_size, last_linear=args.n_embd * 2)
        self.decoder = GPTDecoder(args, embd_pdrop=args.embd_pdrop, resid_pdrop=args.resid_pdrop,
                                  attn_pdrop=args.attn_pdrop, block_size=args.block_size + 1)
        dev = getattr(args, "device", "cpu")
        self.prior = torch.distributions.normal.Normal(torch.zeros(args.n_embd, device=dev), torch.ones(args.n_embd, device=dev))
        self.ns = 2
        self.kl_weight = getattr(args, "kl_start", 1.0)
        self.forward_shuffle_idx, self.backward_shuffle_idx = self.make_idx(5, 53)

    # ---------------------------------------------------------------- inference network (reference :90-105)
    def encode(self, x, nsamples=1):
        """-> z (B, nsamples, nz), KL (B,)"""
        return self.encoder.encode(x, nsamples)

    def encode_stats(self, x):
        """-> mean (B, nz), logvar (B, nz), att"""
        return self.encoder.encode_stats(x)

    def sample_from_inference(self, x, nsamples=1):
        z, _, _ = self.encoder.sample(x, nsamples)
        return z

    # ---------------------------------------------------------------- generator (reference :107-145)
    @torch.no_grad()
    def decode(self, z, strategy, top_k=None, temperature=None):
        """z (B, nsamples, nz) -> (tokens (B, block_size) int64, att);  strategy: "beam" samples with top-k (as the
        reference does), "greedy" / "sample" take the arg-max"""
        was_training = self.training
        if was_training:
            self.eval()
        start = torch.empty((z.size(0), 0), dtype=torch.int64, device=z.device)
        if strategy == "beam":
            out = self.decoder.sample(start, z, steps=self.args.block_size,
                                      temperature=temperature if temperature is not None else 1.0, sample=True,
                                      top_k=top_k if top_k is not None else 100)
        elif strategy in ("greedy", "sample"):
            out = self.decoder.sample(start, z, steps=self.args.block_size, sample=False)
        else:
            raise UnboundLocalError("local variable 'index_sampled' referenced before assignment")   # as the reference
        if was_training:
            self.train()
        return out

    def reconstruct(self, x, decoding_strategy="greedy", K=None):
        """-> tokens, (att_enc, att_dec)   (reference :160-176)"""
        z, _, att_enc = self.encoder.sample(x, nsamples=1)
        rec, att_dec = self.decode(z, decoding_strategy, K)
        return rec, (att_enc, att_dec)

    def loss(self, x, kl_weight, nsamples=1):
        """-> (reconstruction + kl_weight * KL, reconstruction, KL), each (B,)   (reference :179-195)"""
        z, KL = self.encode(x, nsamples)
        reconstruct_err = self.decoder.reconstruct_error(x, z).mean(dim=1)
        return reconstruct_err + kl_weight * KL, reconstruct_err, KL

    # ---------------------------------------------------------------- code order helpers (reference :200-227)
    def make_idx(self, H, W):
        idx = torch.tensor(np.arange(H * W).reshape(H, W).T.ravel())
        return idx, torch.argsort(idx)

    def code_reader(self, x, reverse=False):
        B, L = x.shape
        L_idx = len(self.forward_shuffle_idx)
        if L != L_idx:
            raise NotImplementedError("code_reader: only %d-token clips are supported (got %d)" % (L_idx, L))
        idx = self.backward_shuffle_idx if reverse else self.forward_shuffle_idx
        return x[:, idx.to(x.device)]

    def get_input(self, batch):
        """batch['codes'] (B, 5, 53) -> (B, 265) time-major   (reference :231-245)"""
        x = batch['codes'].to(memory_format=torch.contiguous_format)
        x = torch.flatten(x.permute(0, 2, 1), start_dim=1)
        return x.to(self.args.device)
