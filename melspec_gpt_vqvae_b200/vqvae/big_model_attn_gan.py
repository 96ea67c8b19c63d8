"""Drop-in for the hot-path part of the reference's vqvae/big_model_attn_gan.py.

Same class names, constructor arguments, method signatures and state_dict keys as the
reference, so `model.load_state_dict(torch.load(ckpt))` and the callers
(`extract_codes.get_codes`, `Lit_minGPT.decode_to_img`) work unchanged -- but every tensor
operation runs in libmgv (hand-written sm_100a CUDA behind the C ABI in include/mgv.h):

  VectorQuantizer.forward            reference :19-54   -> mgv_vq_argmin + mgv_vq_finish
  VectorQuantizer.get_codebook_entry reference :56-71   -> mgv_vq_gather
  LitVQVAE.encode                    reference :604-608 -> mgv_vqvae_encode
  LitVQVAE.decode                    reference :610-614 -> mgv_vqvae_decode
  (codes -> mel, used by decode_to_img)                 -> mgv_vqvae_decode_codes

The nn.Module tree below exists to own the parameters under the reference's names; the
packed bf16 copies inside the libmgv handle are derived caches that are refreshed whenever a
parameter changes (load_state_dict, .to(), in-place updates that bump the tensor version; edits through
`.data` need an explicit `refresh_weights()`).  Every libmgv call runs with the parameters' GPU made current
and on that device's torch stream; the handle is rebuilt when the module moves to another GPU.  Inference only (the VQVAE
is never trained in the reference repo: README.md:16); no CPU fallback.
"""
import ctypes

import torch
import torch.nn as nn

from .. import _lib

try:  # the reference derives LitVQVAE from pl.LightningModule (:538); keep that when PL is installed
    import pytorch_lightning as pl
    _LitBase = pl.LightningModule
except Exception:  # pragma: no cover - PL is absent in the build image
    _LitBase = nn.Module

# module-level architecture constants, reference :521-530
double_z = False
z_channels = 256
resolution = 848
in_channels = 1
out_ch = 1
ch = 128
ch_mult = [1, 1, 2, 2, 4]
num_res_blocks = 2
attn_resolutions = [53]
dropout = 0.0

_HOT_PATH_ONLY = ("this submodule only owns parameters; the computation runs fused inside libmgv -- call "
                  "LitVQVAE.encode / LitVQVAE.decode / VectorQuantizer.forward instead")


def _params_signature(module, owner=None):
    """Changes whenever any parameter/buffer is modified in place or moved (data pointer, version counter, device).

    Walking `module.parameters()` costs ~1.3 ms for the VQVAE (GPU idle meanwhile), so the tensor list is cached on
    `owner` (a `_SigCacheMixin` module) and rebuilt after `_apply` (.to / .cuda / .half ...), `load_state_dict` and
    `refresh_weights`; re-binding an attribute to a NEW nn.Parameter object needs `refresh_weights()`."""
    ts = None
    cache = owner.__dict__.get("_mgv_sig_tensors") if owner is not None else None
    if cache is not None:
        ts = cache.get(id(module))
    if ts is None:
        ts = list(module.parameters()) + list(module.buffers())
        if cache is not None:
            cache[id(module)] = ts
    return hash(tuple([(t.data_ptr(), t._version, t.device.index if t.is_cuda else -1) for t in ts]))


class _SigCacheMixin:
    """Invalidation points of the cached tensor list `_params_signature` uses."""

    def _sig_cache_reset(self):
        self.__dict__["_mgv_sig_tensors"] = {}

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._sig_cache_reset()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._sig_cache_reset()
        return out


class VectorQuantizer(nn.Module):
    """reference :8-71"""

    def __init__(self, num_embeddings, embedding_dim, commitment_cost):
        super().__init__()
        self._embedding_dim = embedding_dim
        self._num_embeddings = num_embeddings
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.uniform_(-1 / self._num_embeddings, 1 / self._num_embeddings)   # :16
        self._commitment_cost = commitment_cost

    def _codebook(self, like):
        w = self._embedding.weight.detach()
        if not w.is_cuda:
            raise RuntimeError("VectorQuantizer: parameters are on %s; libmgv has no CPU path -- move the module "
                               "to a B200 (model.to('cuda'))" % w.device)
        if w.dtype != torch.float32:
            raise RuntimeError("VectorQuantizer: codebook must be float32")
        return w.contiguous()

    @torch.no_grad()
    def encoding_indices(self, inputs):
        """Index part of forward only (what extract_codes needs): (N,) int64, (b,h,w) order."""
        cb = self._codebook(inputs)
        z = self._check_input(inputs)
        B, C = z.shape[0], z.shape[1]
        HW = z.shape[2] * z.shape[3]
        with torch.cuda.device(z.device):
            idx = torch.empty(B * HW, dtype=torch.int64, device=z.device)
            _lib.check(_lib.load().mgv_vq_argmin(_lib.ptr(z), _lib.ptr(cb), B, C, HW, self._num_embeddings, _lib.ptr(idx),
                                                 _lib.ptr(self._dmin(B * HW, z.device)), _lib.stream_ptr(z.device)), "mgv_vq_argmin")
        return idx

    def _dmin(self, n, device):
        """running-minimum buffer of the multi-pass search (codebooks beyond 128 codes); caller-owned so that concurrent
        streams / devices never share scratch memory"""
        return torch.empty(max(n, 1), dtype=torch.float32, device=device) if self._num_embeddings > 128 else None

    def _check_input(self, inputs):
        if inputs.dim() != 4 or inputs.shape[1] != self._embedding_dim:
            raise RuntimeError("VectorQuantizer: expected BCHW input with C=%d, got %s" %
                               (self._embedding_dim, tuple(inputs.shape)))
        if not inputs.is_cuda:
            raise RuntimeError("VectorQuantizer: input is on %s; libmgv has no CPU path" % inputs.device)
        z = inputs.detach()
        if z.dtype != torch.float32:
            z = z.float()
        if z.device != self._embedding.weight.device:
            raise RuntimeError("VectorQuantizer: input is on %s but the codebook is on %s" % (z.device, self._embedding.weight.device))
        return z.contiguous()

    @torch.no_grad()
    def forward(self, inputs):
        """-> (loss, quantized BCHW, (perplexity, encodings (N,K) fp32 one-hot, encoding_indices (N,1) int64))"""
        cb = self._codebook(inputs)
        z = self._check_input(inputs)
        B, C, H, W = z.shape
        HW = H * W
        K = self._num_embeddings
        dev = z.device
        L = _lib.load()
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            idx = torch.empty(B * HW, dtype=torch.int64, device=dev)
            _lib.check(L.mgv_vq_argmin(_lib.ptr(z), _lib.ptr(cb), B, C, HW, K, _lib.ptr(idx), _lib.ptr(self._dmin(B * HW, dev)), st),
                       "mgv_vq_argmin")
            quantized = torch.empty_like(z)
            encodings = torch.empty(B * HW, K, dtype=torch.float32, device=dev)
            scalars = torch.zeros(2, dtype=torch.float32, device=dev)
            if B * HW == 0:     # empty batch: mse / perplexity of nothing (the reference returns nan / 1.0)
                scalars[0] = float("nan")
                scalars[1] = 1.0
                return scalars[0], quantized, (scalars[1], encodings, idx.unsqueeze(1))
            ws = torch.empty((8 + 4 * K + 7) // 8, dtype=torch.float64, device=dev)
            _lib.check(L.mgv_vq_finish(_lib.ptr(z), _lib.ptr(cb), _lib.ptr(idx), B, C, HW, K, float(self._commitment_cost),
                                       _lib.ptr(quantized), _lib.ptr(encodings), ctypes.c_void_p(scalars.data_ptr()),
                                       ctypes.c_void_p(scalars.data_ptr() + 4), _lib.ptr(ws), st), "mgv_vq_finish")
        return scalars[0], quantized, (scalars[1], encodings, idx.unsqueeze(1))

    @torch.no_grad()
    def get_codebook_entry(self, indices, shape):
        """shape = (batch, height, width, channel) or None (reference :56-71)."""
        cb = self._codebook(indices)
        if not indices.is_cuda:
            raise RuntimeError("get_codebook_entry: indices are on %s; libmgv has no CPU path" % indices.device)
        idx = indices.reshape(-1).to(torch.int64).contiguous()
        n = idx.numel()
        C = self._embedding_dim
        L = _lib.load()
        flag = torch.zeros(1, dtype=torch.int32, device=idx.device)
        if shape is not None:
            Bq, Hq, Wq, Cq = shape
            if Cq != C or Bq * Hq * Wq != n:
                raise RuntimeError("shape '%s' is invalid for input of size %d" % (list(shape), n * C))
            out = torch.empty(Bq, C, Hq, Wq, dtype=torch.float32, device=idx.device)
            hw = Hq * Wq
        else:
            out = torch.empty(n, C, dtype=torch.float32, device=idx.device)
            hw = 0
        with torch.cuda.device(idx.device):
            _lib.check(L.mgv_vq_gather(_lib.ptr(idx), _lib.ptr(cb), n, C, hw, self._num_embeddings, _lib.ptr(out),
                                       _lib.ptr(flag), _lib.stream_ptr(idx.device)), "mgv_vq_gather")
        if int(flag.item()) != 0:
            raise RuntimeError("index out of range in get_codebook_entry (num_embeddings=%d)" % self._num_embeddings)
        return out


# ----------------------------------------------------------------------------------------------
# Parameter containers with the reference's attribute names (state_dict compatibility).
# ----------------------------------------------------------------------------------------------
def Normalize(in_channels):
    return torch.nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)   # reference :139-140


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError("%s: %s" % (type(self).__name__, _HOT_PATH_ONLY))


class ResnetBlock(_Container):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = Normalize(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(_Container):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)


class Downsample(_Container):
    def __init__(self, in_channels, with_conv=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)


class Upsample(_Container):
    def __init__(self, in_channels, with_conv=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Encoder(_Container):
    """reference :190-282 (parameters only)."""

    def __init__(self):
        super().__init__()
        nres = len(ch_mult)
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        in_ch_mult = (1,) + tuple(ch_mult)
        curr_res = resolution
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(nres):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout, temb_channels=0))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != nres - 1:
                down.downsample = Downsample(block_in)
                curr_res //= 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout, temb_channels=0)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout, temb_channels=0)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, 1, 1)


class Decoder(_Container):
    """reference :291-392 (parameters only)."""

    def __init__(self):
        super().__init__()
        nres = len(ch_mult)
        block_in = ch * ch_mult[nres - 1]
        curr_res = resolution // 2 ** (nres - 1)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout, temb_channels=0)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout, temb_channels=0)
        self.up = nn.ModuleList()
        for i_level in reversed(range(nres)):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout, temb_channels=0))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)


class NLayerDiscriminator(_Container):
    """State container for the PatchGAN critic (reference :465-514): it is part of the reference's
    checkpoints, so its keys must exist for a strict load_state_dict, but it is never on the
    inference path (out of scope)."""

    def __init__(self, input_nc=1, ndf=64, n_layers=3, use_actnorm=False):
        super().__init__()
        layers = [nn.Conv2d(input_nc, ndf, 4, 2, 1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers + 1):
            prev, mult = mult, min(2 ** n, 8)
            stride = 2 if n < n_layers else 1
            layers += [nn.Conv2d(ndf * prev, ndf * mult, 4, stride, 1, bias=False), nn.BatchNorm2d(ndf * mult),
                       nn.LeakyReLU(0.2, True)]
        layers += [nn.Conv2d(ndf * mult, 1, 4, 1, 1)]
        self.main = nn.Sequential(*layers)


class LitVQVAE(_SigCacheMixin, _LitBase):
    """reference :538-614.  encode / decode / _vq_vae are the hot path; the GAN training methods
    (loss, training_step, configure_optimizers ...) are out of scope and not provided."""

    def __init__(self, num_embeddings, embedding_dim, commitment_cost=0.25, disc_start=2001, codebook_weight=1.0,
                 disc_num_layers=3, disc_in_channels=1, disc_factor=1.0, disc_weight=1.0, use_actnorm=False,
                 disc_conditional=False, disc_ndf=64, min_adapt_weight=0.0, max_adapt_weight=1e4, learning_rate=1e-3):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = int(embedding_dim)
        self._encoder = Encoder()
        self._vq_vae = VectorQuantizer(num_embeddings, self.embedding_dim, commitment_cost)
        self._decoder = Decoder()
        self.quant_conv = nn.Conv2d(z_channels, self.embedding_dim, 1)
        self.post_quant_conv = nn.Conv2d(self.embedding_dim, z_channels, 1)
        self.counts = [0 for _ in range(self.num_embeddings)]
        self.learning_rate = learning_rate
        self.codebook_weight = codebook_weight
        self.discriminator = NLayerDiscriminator(input_nc=disc_in_channels, n_layers=disc_num_layers,
                                                 use_actnorm=use_actnorm, ndf=disc_ndf)
        self.discriminator_iter_start = disc_start * 2
        self.disc_factor = disc_factor
        self.discriminator_weight = disc_weight
        self.disc_conditional = disc_conditional
        self.min_adapt_weight = min_adapt_weight
        self.max_adapt_weight = max_adapt_weight
        self._mgv_handle = None
        self._mgv_sig = None
        self._mgv_dev = None
        self._sig_cache_reset()

    # ---------------------------------------------------------------- libmgv handle
    def _hot_modules(self):
        return (self._encoder, self._decoder, self._vq_vae, self.quant_conv, self.post_quant_conv)

    def _handle(self):
        """libmgv handle for the device that holds the parameters (call with that device current: `_on_device`)."""
        p = self.quant_conv.weight
        if not p.is_cuda:
            raise RuntimeError("LitVQVAE: parameters are on %s; libmgv has no CPU path -- call .to('cuda')" % p.device)
        L = _lib.load()
        sig = hash(tuple(_params_signature(m, self) for m in self._hot_modules()))
        if self._mgv_handle is not None and self._mgv_dev != p.device:
            self._release_handle()                     # the module moved to another GPU
        if self._mgv_handle is None:
            h = ctypes.c_void_p()
            _lib.check(L.mgv_vqvae_create(int(self.num_embeddings), self.embedding_dim, ctypes.byref(h)), "mgv_vqvae_create")
            self._mgv_handle = h
            self._mgv_dev = p.device
            self._mgv_sig = None
        if sig != self._mgv_sig:
            st = _lib.stream_ptr(p.device)
            for prefix, mod in (("_encoder.", self._encoder), ("_decoder.", self._decoder), ("_vq_vae.", self._vq_vae),
                                ("quant_conv.", self.quant_conv), ("post_quant_conv.", self.post_quant_conv)):
                for k, t in mod.state_dict().items():
                    t32 = t.detach().to(torch.float32).contiguous()
                    _lib.check(L.mgv_vqvae_load_weight(self._mgv_handle, (prefix + k).encode(), _lib.ptr(t32),
                                                       t32.numel(), st), "mgv_vqvae_load_weight(%s%s)" % (prefix, k))
            torch.cuda.current_stream(p.device).synchronize()
            self._mgv_sig = sig
        return self._mgv_handle

    def _on_device(self):
        p = self.quant_conv.weight
        if not p.is_cuda:
            raise RuntimeError("LitVQVAE: parameters are on %s; libmgv has no CPU path -- call .to('cuda')" % p.device)
        return torch.cuda.device(p.device)

    def refresh_weights(self):
        """Re-pack the bf16 weight copies inside libmgv on the next call (needed only after edits through `.data`, which
        bypass the version counter the automatic check relies on, or after re-binding an attribute to a new nn.Parameter)."""
        self._mgv_sig = None
        self._sig_cache_reset()

    def _release_handle(self):
        if getattr(self, "_mgv_handle", None) is not None:
            try:
                with torch.cuda.device(self._mgv_dev):
                    _lib.load().mgv_vqvae_destroy(self._mgv_handle)
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

    @staticmethod
    def _f32_cuda(x, what):
        if not x.is_cuda:
            raise RuntimeError("%s: input is on %s; libmgv has no CPU path" % (what, x.device))
        return x.detach().to(torch.float32).contiguous()

    # ---------------------------------------------------------------- hot path
    @torch.no_grad()
    def encode(self, x):
        """mel (B,1,80,848) in [-1,1] -> z (B, embedding_dim, 5, 53)   (reference :604-608)"""
        x = self._f32_cuda(x, "LitVQVAE.encode")
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 80, resolution):
            raise RuntimeError("LitVQVAE.encode: expected (B,1,80,%d), got %s" % (resolution, tuple(x.shape)))
        B = x.shape[0]
        with self._on_device():
            z = torch.empty(B, self.embedding_dim, 5, 53, dtype=torch.float32, device=x.device)
            _lib.check(_lib.load().mgv_vqvae_encode(self._handle(), _lib.ptr(x), B, _lib.ptr(z), _lib.stream_ptr(x.device)),
                       "mgv_vqvae_encode")
        return z

    @torch.no_grad()
    def decode(self, quant):
        """z_q (B, embedding_dim, 5, 53) -> mel (B,1,80,848)   (reference :610-614)"""
        q = self._f32_cuda(quant, "LitVQVAE.decode")
        if q.dim() != 4 or tuple(q.shape[1:]) != (self.embedding_dim, 5, 53):
            raise RuntimeError("LitVQVAE.decode: expected (B,%d,5,53), got %s" % (self.embedding_dim, tuple(q.shape)))
        B = q.shape[0]
        with self._on_device():
            mel = torch.empty(B, 1, 80, resolution, dtype=torch.float32, device=q.device)
            _lib.check(_lib.load().mgv_vqvae_decode(self._handle(), _lib.ptr(q), B, _lib.ptr(mel), _lib.stream_ptr(q.device)),
                       "mgv_vqvae_decode")
        return mel

    @torch.no_grad()
    def decode_codes(self, index_row_major):
        """(B, 265) int64 code grid (row-major 5x53) -> mel.  Equals
        decode(_vq_vae.get_codebook_entry(index.reshape(-1), (B,5,53,C))) with the gather and
        post_quant_conv fused into one table lookup."""
        idx = index_row_major
        if not idx.is_cuda:
            raise RuntimeError("LitVQVAE.decode_codes: indices are on %s; libmgv has no CPU path" % idx.device)
        idx = idx.to(torch.int64).reshape(-1, 265).contiguous()
        B = idx.shape[0]
        with self._on_device():
            mel = torch.empty(B, 1, 80, resolution, dtype=torch.float32, device=idx.device)
            _lib.check(_lib.load().mgv_vqvae_decode_codes(self._handle(), _lib.ptr(idx), B, _lib.ptr(mel),
                                                          _lib.stream_ptr(idx.device)), "mgv_vqvae_decode_codes")
        return mel

    @torch.no_grad()
    def forward(self, x):
        """reference :622-634 (eval): returns (loss, x_recon, info)."""
        z = self.encode(x)
        loss, quantized, info = self._vq_vae(z)
        x_recon = self.decode(quantized)
        if not self.training:
            binc = torch.bincount(info[2].reshape(-1), minlength=self.num_embeddings).tolist()
            self.counts = [a + b for a, b in zip(binc, self.counts)]
        return loss, x_recon, info

    def last_launches(self):
        with self._on_device():
            return int(_lib.load().mgv_vqvae_last_launches(self._handle()))
