"""ORACLE -- test infrastructure, not product code.  Only tests/, __graft_entry__.smoke() and bench.py's reference legs
may use it.

fp32 torch-CPU functional restatement of the reference's MelGAN Generator (vocoder/modules.py:23-36 ResnetBlock,
:38-80 Generator), working straight on a state_dict with the reference's keys (weight_g / weight_v / bias).  Pinned
against the unmodified reference class by tests/golden/melgan_small.npz (tests/golden/make_golden_melgan.py)."""
import torch
import torch.nn.functional as F

RATIOS = (8, 8, 2, 2)


def _w(sd, prefix):
    """torch.nn.utils.weight_norm with dim=0 (reference :17-21): w = g * v / |v|, the norm over all dims but the first"""
    v, g = sd[prefix + ".weight_v"].float(), sd[prefix + ".weight_g"].float()
    return v * (g / v.flatten(1).norm(dim=1).reshape(-1, 1, 1)), sd[prefix + ".bias"].float()


def generator_forward(sd, x, n_residual_layers=3):
    """x (B, n_mel, T) -> (B, 1, 256 T)   (Generator.forward :79-80 over the nn.Sequential built at :46-77)"""
    w, b = _w(sd, "model.1")
    h = F.conv1d(F.pad(x.float(), (3, 3), mode="reflect"), w, b)                     # :46-49
    i = 2
    for r in RATIOS:
        w, b = _w(sd, "model.%d" % (i + 1))
        h = F.conv_transpose1d(F.leaky_relu(h, 0.2), w, b, stride=r, padding=r // 2 + r % 2, output_padding=r % 2)   # :53-63
        for j in range(n_residual_layers):                                          # :65-66, ResnetBlock :23-36
            p = "model.%d" % (i + 2 + j)
            d = 3 ** j
            w3, b3 = _w(sd, p + ".block.2")
            w1, b1 = _w(sd, p + ".block.4")
            ws, bs = _w(sd, p + ".shortcut")
            y = F.conv1d(F.pad(F.leaky_relu(h, 0.2), (d, d), mode="reflect"), w3, b3, dilation=d)
            y = F.conv1d(F.leaky_relu(y, 0.2), w1, b1)
            h = F.conv1d(h, ws, bs) + y
        i += 2 + n_residual_layers
    w, b = _w(sd, "model.%d" % (i + 2))
    return torch.tanh(F.conv1d(F.pad(F.leaky_relu(h, 0.2), (3, 3), mode="reflect"), w, b))   # :70-75
