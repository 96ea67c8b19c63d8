"""ORACLE -- test infrastructure, not product code.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

CPU restatement of the reference quantiser (karchkha/MelSpec_GPT_VQVAE):
  VectorQuantizer.forward            vqvae/big_model_attn_gan.py:19-54
  VectorQuantizer.get_codebook_entry vqvae/big_model_attn_gan.py:56-71

Two restatements live here:
  * `forward_numpy` / `get_codebook_entry_numpy`: the reference's formula line by line in
    numpy fp32 (summation order = numpy/BLAS's, like the reference's = torch/MKL's);
  * `argmin_exact` / `distances_exact` (vq_oracle.c through ctypes): the same formula with
    the summation order FIXED (sequential fmaf over channels), which the CUDA kernel
    reproduces bit for bit.
Parity pin: tests/golden/vq_*.npz hold outputs of the UNMODIFIED reference module on
seeded inputs (tests/golden/make_golden.py); tests/test_oracle_cpu.py checks both
restatements against them (indices equal except classified near-ties).
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvq_oracle.so")
_lib = None


def build(force=False):
    """gcc-compile vq_oracle.c (also called by __graft_entry__.build)."""
    src = os.path.join(_HERE, "vq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", _SO, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        for name in ("vq_argmin_oracle", "vq_distances_oracle", "vq_gather_oracle"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def argmin_exact(z_bchw: np.ndarray, codebook: np.ndarray, threads: int = 0):
    """z (B, D, H, W) or (B, D, HW) fp32 -> (idx int64 (B*HW,), dmin fp32 (B*HW,)); fixed summation order."""
    lib = _load()
    z = np.ascontiguousarray(z_bchw, dtype=np.float32)
    B, D = z.shape[0], z.shape[1]
    HW = int(np.prod(z.shape[2:]))
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    K = cb.shape[0]
    assert cb.shape[1] == D and K <= 4096
    idx = np.empty(B * HW, dtype=np.int64)
    dmin = np.empty(B * HW, dtype=np.float32)
    if B == 0:
        return idx, dmin
    threads = threads or min(os.cpu_count() or 1, B)
    z3 = z.reshape(B, D, HW)

    def run(lo, hi):
        lo, hi = int(lo), int(hi)
        lib.vq_argmin_oracle(_p(z3[lo:hi]), _p(cb), hi - lo, int(D), HW, int(K), _p(idx[lo * HW:hi * HW]),
                             _p(dmin[lo * HW:hi * HW]))

    bounds = np.linspace(0, B, threads + 1).astype(int)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda i: run(bounds[i], bounds[i + 1]) if bounds[i + 1] > bounds[i] else None, range(threads)))
    return idx, dmin


def distances_exact(z_bchw: np.ndarray, codebook: np.ndarray) -> np.ndarray:
    """(B*HW, K) fp32 distances in the fixed summation order."""
    lib = _load()
    z = np.ascontiguousarray(z_bchw, dtype=np.float32)
    B, D = z.shape[0], z.shape[1]
    HW = int(np.prod(z.shape[2:]))
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    dist = np.empty((B * HW, cb.shape[0]), dtype=np.float32)
    lib.vq_distances_oracle(_p(z.reshape(B, D, HW)), _p(cb), int(B), int(D), HW, int(cb.shape[0]), _p(dist))
    return dist


def forward_numpy(inputs_bchw: np.ndarray, codebook: np.ndarray, commitment_cost: float, indices=None):
    """Line-by-line numpy fp32 restatement of VectorQuantizer.forward (:19-54).
    Returns (loss, quantized_bchw, (perplexity, encodings, encoding_indices (N,1) int64)).
    `indices`: optionally force the encoding indices (to compare the remaining outputs on
    identical indices)."""
    x = np.ascontiguousarray(np.transpose(inputs_bchw.astype(np.float32), (0, 2, 3, 1)))      # :21
    input_shape = x.shape
    D = codebook.shape[1]
    K = codebook.shape[0]
    flat = x.reshape(-1, D)                                                                  # :25
    cb = codebook.astype(np.float32)
    if indices is None:
        distances = ((flat ** 2).sum(axis=1, keepdims=True, dtype=np.float32)                # :28
                     + (cb ** 2).sum(axis=1, dtype=np.float32)                               # :29
                     - np.float32(2) * (flat @ cb.T))                                        # :30
        enc_idx = np.argmin(distances, axis=1).astype(np.int64)[:, None]                     # :33
    else:
        enc_idx = np.asarray(indices, dtype=np.int64).reshape(-1, 1)
    encodings = np.zeros((enc_idx.shape[0], K), dtype=np.float32)                             # :36
    encodings[np.arange(enc_idx.shape[0]), enc_idx[:, 0]] = 1.0                               # :37
    quantized = cb[enc_idx[:, 0]].reshape(input_shape)                                        # :40 (one-hot @ W == gather)
    mse = np.mean((quantized.astype(np.float64) - x.astype(np.float64)) ** 2)
    e_latent = np.float32(mse)                                                                # :43
    q_latent = np.float32(mse)                                                                # :44
    loss = np.float32(q_latent + np.float32(commitment_cost) * e_latent)                      # :45
    quantized_st = x + (quantized - x)                                                        # :49
    avg_probs = encodings.mean(axis=0, dtype=np.float32)                                      # :50
    perplexity = np.exp(-np.sum(avg_probs * np.log(avg_probs + np.float32(1e-10)), dtype=np.float32))  # :51
    q_bchw = np.ascontiguousarray(np.transpose(quantized_st, (0, 3, 1, 2)))                   # :54
    return loss, q_bchw, (np.float32(perplexity), encodings, enc_idx)


def get_codebook_entry_numpy(indices: np.ndarray, codebook: np.ndarray, shape):
    """VectorQuantizer.get_codebook_entry (:56-71): shape = (B,H,W,C) or None."""
    z_q = codebook.astype(np.float32)[np.asarray(indices, dtype=np.int64)]                    # :59-63
    if shape is not None:
        z_q = z_q.reshape(shape)                                                              # :66
        z_q = np.ascontiguousarray(np.transpose(z_q, (0, 3, 1, 2)))                           # :69
    return z_q


def classify_mismatches(dist: np.ndarray, idx_a: np.ndarray, idx_b: np.ndarray, ulps: int = 8):
    """For rows where two argmin results differ: is the gap between the two chosen distances
    within `ulps` units in the last place of the distance magnitude (a tie) or a real error?
    Returns dict(n_mismatch, n_tie, n_real, max_gap_ulps)."""
    idx_a = np.asarray(idx_a).reshape(-1)
    idx_b = np.asarray(idx_b).reshape(-1)
    rows = np.nonzero(idx_a != idx_b)[0]
    n_tie = n_real = 0
    max_gap = 0.0
    for r in rows:
        da, db = dist[r, idx_a[r]], dist[r, idx_b[r]]
        ulp = np.spacing(np.float32(max(abs(da), abs(db), np.finfo(np.float32).tiny)))
        gap = abs(float(da) - float(db)) / float(ulp)
        max_gap = max(max_gap, gap)
        if gap <= ulps:
            n_tie += 1
        else:
            n_real += 1
    return dict(n_mismatch=int(rows.size), n_tie=n_tie, n_real=n_real, max_gap_ulps=max_gap)
