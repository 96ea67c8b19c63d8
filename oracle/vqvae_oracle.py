"""ORACLE -- test infrastructure, not product code.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

CPU fp32 restatement (functional torch ops on a plain state_dict) of the reference VQVAE
encoder / decoder of karchkha/MelSpec_GPT_VQVAE, vqvae/big_model_attn_gan.py:
  ResnetBlock.forward :114-135   Normalize :139-140 (GroupNorm 32, eps 1e-6)
  Downsample.forward :156-162    nonlinearity :164-166 (swish)
  Upsample.forward :182-186      Encoder.forward :254-282
  Decoder.forward :361-392       AttnBlock.forward :425-450
  LitVQVAE.encode :604-608       LitVQVAE.decode :610-614
  module constants :521-530 (ch 128, ch_mult [1,1,2,2,4], 2 res blocks, attn at 53)
Parity pin: tests/golden/vqvae_*.npz (outputs of the UNMODIFIED reference on seeded
synthetic weights, tests/golden/make_golden.py), checked by tests/test_oracle_cpu.py.
"""
import torch
import torch.nn.functional as F

CH = 128
CH_MULT = [1, 1, 2, 2, 4]
NUM_RES_BLOCKS = 2
ATTN_RESOLUTIONS = [53]
RESOLUTION = 848
Z_CHANNELS = 256


def swish(x):
    return x * torch.sigmoid(x)                                                      # :164-166


def norm(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)         # :139-140


def conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(sd, p, x):
    h = conv(sd, p + ".conv1", swish(norm(sd, p + ".norm1", x)))                     # :117-119
    h = conv(sd, p + ".conv2", swish(norm(sd, p + ".norm2", h)))                     # :124-127 (dropout 0)
    if (p + ".nin_shortcut.weight") in sd:
        x = conv(sd, p + ".nin_shortcut", x, padding=0)                              # :133
    return x + h                                                                     # :135


def attn_block(sd, p, x):
    h_ = norm(sd, p + ".norm", x)                                                    # :428
    q = conv(sd, p + ".q", h_, padding=0)
    k = conv(sd, p + ".k", h_, padding=0)
    v = conv(sd, p + ".v", h_, padding=0)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)                                      # :435-436
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))                                        # :438-439
    w_ = F.softmax(w_, dim=2)                                                        # :440
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)                       # :444-446
    h_ = conv(sd, p + ".proj_out", h_, padding=0)                                    # :448
    return x + h_


@torch.no_grad()
def encoder(sd, x, prefix="_encoder"):
    """Encoder.forward (:254-282)."""
    num_res = len(CH_MULT)
    h = conv(sd, prefix + ".conv_in", x)
    curr_res = RESOLUTION
    for i_level in range(num_res):
        has_attn = curr_res in ATTN_RESOLUTIONS
        for i_block in range(NUM_RES_BLOCKS):
            h = resnet_block(sd, "%s.down.%d.block.%d" % (prefix, i_level, i_block), h)
            if has_attn:
                h = attn_block(sd, "%s.down.%d.attn.%d" % (prefix, i_level, i_block), h)
        if i_level != num_res - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)                      # :158
            h = conv(sd, "%s.down.%d.downsample.conv" % (prefix, i_level), h, stride=2, padding=0)
            curr_res //= 2
    h = resnet_block(sd, prefix + ".mid.block_1", h)
    h = attn_block(sd, prefix + ".mid.attn_1", h)
    h = resnet_block(sd, prefix + ".mid.block_2", h)
    h = swish(norm(sd, prefix + ".norm_out", h))
    return conv(sd, prefix + ".conv_out", h)


@torch.no_grad()
def decoder(sd, z, prefix="_decoder"):
    """Decoder.forward (:361-392)."""
    num_res = len(CH_MULT)
    h = conv(sd, prefix + ".conv_in", z)
    h = resnet_block(sd, prefix + ".mid.block_1", h)
    h = attn_block(sd, prefix + ".mid.attn_1", h)
    h = resnet_block(sd, prefix + ".mid.block_2", h)
    curr_res = RESOLUTION // 2 ** (num_res - 1)
    for i_level in reversed(range(num_res)):
        has_attn = curr_res in ATTN_RESOLUTIONS
        for i_block in range(NUM_RES_BLOCKS + 1):
            h = resnet_block(sd, "%s.up.%d.block.%d" % (prefix, i_level, i_block), h)
            if has_attn:
                h = attn_block(sd, "%s.up.%d.attn.%d" % (prefix, i_level, i_block), h)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")                    # :183
            h = conv(sd, "%s.up.%d.upsample.conv" % (prefix, i_level), h)
            curr_res *= 2
    h = swish(norm(sd, prefix + ".norm_out", h))
    return conv(sd, prefix + ".conv_out", h)


@torch.no_grad()
def encode(sd, x):
    """LitVQVAE.encode (:604-608)."""
    return conv(sd, "quant_conv", encoder(sd, x), padding=0)


@torch.no_grad()
def decode(sd, quant):
    """LitVQVAE.decode (:610-614)."""
    return decoder(sd, conv(sd, "post_quant_conv", quant, padding=0))


@torch.no_grad()
def decode_codes(sd, index_row_major, B, H=5, W=53):
    """get_codebook_entry (:56-71) + decode: index (B*H*W,) in (b,h,w) order."""
    cb = sd["_vq_vae._embedding.weight"]
    z_q = cb[index_row_major.reshape(-1)].view(B, H, W, cb.shape[1]).permute(0, 3, 1, 2).contiguous()
    return decode(sd, z_q)
