"""Drop-in for the reference's vocoder/modules.py: the MelGAN `Generator` the logging callbacks use to turn decoded
mel spectrograms into audio (the step after the token path's end).

    Generator(input_size, ngf, n_residual_layers)   reference vocoder/modules.py:38-80   -> mgv_melgan_forward
    callers: callbacks/GPT_callbacks.py:66-79 (load_vocoder), :93-105 (_log_rec_audio)

Same constructor, same `state_dict` keys (`model.<i>.weight_g / weight_v / bias`, ResnetBlocks under
`model.<i>.block.{2,4}` and `model.<i>.shortcut`: a `best_netG.pt` of the reference loads unchanged), same
`hop_length`, same input / output tensors.  The weight norm (torch.nn.utils.weight_norm, reference :17-21) is applied
when the weights are packed for libmgv; the computation is fp32 CUDA (csrc/melgan.cu).  Inference only; no CPU path.
"""
import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib

RATIOS = (8, 8, 2, 2)          # reference :41


class _WNParams(nn.Module):
    """Parameters of one weight-normed convolution under the reference's names (weight_g, weight_v, bias)."""

    def __init__(self, shape, bias_len):
        super().__init__()
        v = torch.empty(*shape)
        fan_in = shape[1] * shape[2]
        nn.init.kaiming_uniform_(v, a=math.sqrt(5))             # nn.Conv1d / nn.ConvTranspose1d default initialiser
        self.weight_v = nn.Parameter(v)
        self.weight_g = nn.Parameter(v.detach().flatten(1).norm(dim=1).reshape(-1, 1, 1).clone())
        bound = 1.0 / math.sqrt(fan_in)
        self.bias = nn.Parameter(torch.empty(bias_len).uniform_(-bound, bound))

    def effective_weight(self):
        """g * v / |v| with the norm over every dimension but the first (weight_norm's default dim=0)."""
        v = self.weight_v.detach().to(torch.float32)
        g = self.weight_g.detach().to(torch.float32)
        return (v * (g / v.flatten(1).norm(dim=1).reshape(-1, 1, 1))).contiguous()

    def forward(self, *a, **k):
        raise NotImplementedError("this submodule only owns parameters; call Generator.forward")


class _Conv(_WNParams):
    def __init__(self, cin, cout, k):
        super().__init__((cout, cin, k), cout)


class _ConvT(_WNParams):
    def __init__(self, cin, cout, k):
        super().__init__((cin, cout, k), cout)       # ConvTranspose1d stores (in, out, k): the norm runs per INPUT channel


class _Res(nn.Module):
    """ResnetBlock (reference :23-36): parameters of block.2 (k=3, dilated), block.4 (1x1) and shortcut (1x1)."""

    def __init__(self, dim):
        super().__init__()
        self.block = nn.ModuleDict({"2": _Conv(dim, dim, 3), "4": _Conv(dim, dim, 1)})
        self.shortcut = _Conv(dim, dim, 1)

    def forward(self, *a, **k):
        raise NotImplementedError("this submodule only owns parameters; call Generator.forward")


class Generator(nn.Module):
    def __init__(self, input_size, ngf, n_residual_layers):
        super().__init__()
        self.input_size, self.ngf, self.n_residual_layers = int(input_size), int(ngf), int(n_residual_layers)
        self.hop_length = int(np.prod(RATIOS))
        mult = 2 ** len(RATIOS)
        layers = {"1": _Conv(input_size, mult * ngf, 7)}        # index 0 is the ReflectionPad1d
        i = 2
        for r in RATIOS:
            layers[str(i + 1)] = _ConvT(mult * ngf, mult * ngf // 2, 2 * r)     # index i is the LeakyReLU
            for j in range(n_residual_layers):
                layers[str(i + 2 + j)] = _Res(mult * ngf // 2)
            i += 2 + n_residual_layers
            mult //= 2
        layers[str(i + 2)] = _Conv(ngf, 1, 7)                   # LeakyReLU, ReflectionPad1d, conv, Tanh
        self.model = nn.ModuleDict(layers)
        self._mgv_handle = None
        self._mgv_dev = None
        self._mgv_sig = None

    # ---- libmgv handle (same life cycle as LitVQVAE._handle)
    def _signature(self):
        return hash(tuple((p.data_ptr(), p._version, p.device.index if p.is_cuda else -1) for p in self.parameters()))

    def _named_convs(self):
        for idx, mod in self.model.items():
            if isinstance(mod, _Res):
                yield "model.%s.block.2" % idx, mod.block["2"]
                yield "model.%s.block.4" % idx, mod.block["4"]
                yield "model.%s.shortcut" % idx, mod.shortcut
            else:
                yield "model.%s" % idx, mod

    def _handle(self):
        p = self.model["1"].weight_v
        if not p.is_cuda:
            raise RuntimeError("Generator: parameters are on %s; libmgv has no CPU path -- call .to('cuda')" % p.device)
        L = _lib.load()
        if self._mgv_handle is not None and self._mgv_dev != p.device:
            self._release_handle()
        if self._mgv_handle is None:
            h = ctypes.c_void_p()
            _lib.check(L.mgv_melgan_create(self.input_size, self.ngf, self.n_residual_layers, ctypes.byref(h)), "mgv_melgan_create")
            self._mgv_handle, self._mgv_dev, self._mgv_sig = h, p.device, None
        sig = self._signature()
        if sig != self._mgv_sig:
            st = _lib.stream_ptr(p.device)
            _lib.check(L.mgv_melgan_reset_biases(self._mgv_handle, st), "mgv_melgan_reset_biases")
            for name, conv in self._named_convs():
                w = conv.effective_weight()
                b = conv.bias.detach().to(torch.float32).contiguous()
                _lib.check(L.mgv_melgan_load_weight(self._mgv_handle, (name + ".weight").encode(), _lib.ptr(w), w.numel(), st),
                           "mgv_melgan_load_weight(%s.weight)" % name)
                _lib.check(L.mgv_melgan_load_weight(self._mgv_handle, (name + ".bias").encode(), _lib.ptr(b), b.numel(), st),
                           "mgv_melgan_load_weight(%s.bias)" % name)
            torch.cuda.current_stream(p.device).synchronize()
            self._mgv_sig = sig
        return self._mgv_handle

    def refresh_weights(self):
        self._mgv_sig = None

    def _release_handle(self):
        if getattr(self, "_mgv_handle", None) is not None:
            try:
                with torch.cuda.device(self._mgv_dev):
                    _lib.load().mgv_melgan_destroy(self._mgv_handle)
            finally:
                self._mgv_handle = None
                self._mgv_sig = None

    def __del__(self):
        try:
            self._release_handle()
        except Exception:
            pass

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_mgv_handle"] = d["_mgv_dev"] = d["_mgv_sig"] = None
        return d

    def last_launches(self):
        return int(_lib.load().mgv_melgan_last_launches(self._mgv_handle)) if self._mgv_handle is not None else 0

    def forward(self, x):
        """x: (B, n_mel, T) float32 on the GPU -> (B, 1, hop_length * T)   (reference :79-80)"""
        if x.dim() != 3 or x.shape[1] != self.input_size:
            raise RuntimeError("Generator: expected (B, %d, T) input, got %s" % (self.input_size, tuple(x.shape)))
        if not x.is_cuda:
            raise RuntimeError("Generator: input is on %s; libmgv has no CPU path" % x.device)
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            pass    # the reference only ever runs the pretrained vocoder under no_grad / eval: nothing is recorded here either
        p = self.model["1"].weight_v
        if x.device != p.device:
            raise RuntimeError("Generator: input on %s, parameters on %s" % (x.device, p.device))
        x = x.detach().to(torch.float32).contiguous()
        B, _, T = x.shape
        with torch.cuda.device(p.device):
            h = self._handle()
            out = torch.empty(B, 1, self.hop_length * T, dtype=torch.float32, device=x.device)
            _lib.check(_lib.load().mgv_melgan_forward(h, _lib.ptr(x), B, T, _lib.ptr(out), _lib.stream_ptr(p.device)),
                       "mgv_melgan_forward")
        return out
