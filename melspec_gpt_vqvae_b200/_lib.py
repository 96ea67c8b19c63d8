"""ctypes binding of libmgv.so (include/mgv.h).  There is no CPU fallback: if the library is
missing, or a call fails, a RuntimeError is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmgv.so")

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
VP = ctypes.c_void_p
I = ctypes.c_int
I64 = ctypes.c_int64
F = ctypes.c_float


class GptConfig(ctypes.Structure):
    _fields_ = [("vocab_size", I), ("block_size", I), ("n_layer", I), ("n_head", I), ("n_embd", I),
                ("class_size", I), ("n_unmasked", I), ("head_out", I)]


# name -> (restype, argtypes); every symbol include/mgv.h declares
SIGNATURES = {
    "mgv_version": (I, []),
    "mgv_last_error": (ctypes.c_char_p, []),
    "mgv_device_check": (I, []),
    "mgv_vq_argmin": (I, [VP, VP, I, I, I, I, VP, VP, VP]),
    "mgv_vq_finish": (I, [VP, VP, VP, I, I, I, I, F, VP, VP, VP, VP, VP, VP]),
    "mgv_vq_gather": (I, [VP, VP, I64, I, I, I, VP, VP, VP]),
    "mgv_gpt_create": (I, [ctypes.POINTER(GptConfig), ctypes.POINTER(VP)]),
    "mgv_gpt_destroy": (I, [VP]),
    "mgv_gpt_load_weight": (I, [VP, ctypes.c_char_p, VP, I64, VP]),
    "mgv_gpt_forward": (I, [VP, VP, I, I, VP, VP, I, VP, VP, VP]),
    "mgv_gpt_generate": (I, [VP, VP, I, I, VP, VP, I, I, F, I, I, ctypes.c_uint64, VP, VP, I, VP]),
    "mgv_gpt_cross_entropy": (I, [VP, VP, VP, I64, I, VP, VP]),
    "mgv_gpt_last_launches": (I64, [VP]),
    "mgv_gpt_set_step_logits": (I, [VP, VP]),
    "mgv_gpt_set_deterministic": (I, [VP, I]),
    "mgv_gpt_train_numel": (I64, [VP]),
    "mgv_gpt_train_layout": (I, [VP, ctypes.c_char_p, ctypes.POINTER(I64), ctypes.POINTER(I64), ctypes.POINTER(I)]),
    "mgv_gpt_train_bind": (I, [VP, VP, VP, VP]),
    "mgv_gpt_train_forward": (I, [VP, VP, I, I, VP, I, VP, F, F, F, ctypes.c_uint64, VP, VP]),
    "mgv_gpt_train_backward": (I, [VP, I, I, VP]),
    "mgv_gpt_train_adamw": (I, [VP, VP, VP, F, F, F, F, F, I64, F, VP]),
    "mgv_test_dropout_mask": (I, [ctypes.c_uint64, ctypes.c_uint, F, I64, VP, VP]),
    "mgv_vqvae_create": (I, [I, I, ctypes.POINTER(VP)]),
    "mgv_vqvae_destroy": (I, [VP]),
    "mgv_vqvae_load_weight": (I, [VP, ctypes.c_char_p, VP, I64, VP]),
    "mgv_vqvae_decode_codes": (I, [VP, VP, I, VP, VP]),
    "mgv_vqvae_decode": (I, [VP, VP, I, VP, VP]),
    "mgv_vqvae_encode": (I, [VP, VP, I, VP, VP]),
    "mgv_vqvae_last_launches": (I64, [VP]),
    "mgv_melgan_create": (I, [I, I, I, ctypes.POINTER(VP)]),
    "mgv_melgan_destroy": (I, [VP]),
    "mgv_melgan_load_weight": (I, [VP, ctypes.c_char_p, VP, I64, VP]),
    "mgv_melgan_reset_biases": (I, [VP, VP]),
    "mgv_melgan_forward": (I, [VP, VP, I, I, VP, VP]),
    "mgv_melgan_last_launches": (I64, [VP]),
    "mgv_test_gemm": (I, [I, VP, VP, I, I, I, I, VP, VP, VP, I, I, VP]),
    "mgv_test_gemm_swapab": (I, [I, VP, VP, I, I, I, I, VP, VP, VP, I, I, VP]),
    "mgv_test_gemm_fold": (I, [I, VP, VP, I, I, I, VP, VP, VP, VP, I, I, VP, I, I, VP]),
    "mgv_test_conv3x3": (I, [I, VP, VP, VP, I, I, I, I, I, I, VP, VP, VP]),
    "mgv_test_attention_prefill": (I, [I, VP, I, I, I, VP, VP, VP]),
    "mgv_test_conv_upsample": (I, [I, VP, VP, VP, I, I, I, I, I, VP, VP, VP]),
}

_lib = None


def load():
    """Load libmgv.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmgv.so not found at %s: build it with `python -m melspec_gpt_vqvae_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().mgv_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "libmgv"):
    if rc != 0:
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, last_error()))


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libmgv needs CUDA tensors (got %s): there is no CPU fallback" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("libmgv needs contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (default: the current device)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
