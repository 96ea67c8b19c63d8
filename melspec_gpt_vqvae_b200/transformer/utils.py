"""Drop-in for the reference's transformer/utils.py helpers used on the GPT-VAE path (safe_log :3-4,
log_sum_exp :6-19).  Host-side scalar statistics; nothing here is a device hot spot."""
import torch


def safe_log(z):
    return torch.log(z + 1e-7)


def log_sum_exp(value, dim=None, keepdim=False):
    """value.exp().sum(dim, keepdim).log(), shifted by the maximum"""
    if dim is None:
        m = torch.max(value)
        return m + torch.log(torch.sum(torch.exp(value - m)))
    m, _ = torch.max(value, dim=dim, keepdim=True)
    out = m + torch.log(torch.sum(torch.exp(value - m), dim=dim, keepdim=True))
    return out if keepdim else out.squeeze(dim)
