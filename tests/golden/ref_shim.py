"""Import shims that let the UNMODIFIED reference (/root/reference) be imported in the
build container, where pytorch_lightning / albumentations are absent and HuggingFace
`datasets` shadows the reference's `datasets/` directory.

Test infrastructure only.  Used by `make_golden.py` (fixture generation) and by the
optional `tests/test_against_reference.py` cross-checks, which skip when /root/reference
is absent (it does not exist on the GPU box).  Nothing in the product imports this.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MGV_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "transformer", "minGPT.py"))


def install():
    """Install the shims and put the reference first on sys.path.  Idempotent."""
    import torch

    if "pytorch_lightning" not in sys.modules or not hasattr(sys.modules["pytorch_lightning"], "_mgv_shim"):
        pl = types.ModuleType("pytorch_lightning")
        pl._mgv_shim = True

        class LightningModule(torch.nn.Module):
            def log(self, *a, **k):
                pass

            def print(self, *a, **k):
                print(*a, **k)

            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

        class LightningDataModule:
            def __init__(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        sys.modules["pytorch_lightning"] = pl

    if "albumentations" not in sys.modules:
        alb = types.ModuleType("albumentations")

        class CenterCrop:
            def __init__(self, h, w):
                self.h, self.w = h, w

            def __call__(self, image):
                H, W = image.shape[:2]
                y0 = (H - self.h) // 2
                x0 = (W - self.w) // 2
                return {"image": image[y0:y0 + self.h, x0:x0 + self.w]}

        class RandomCrop(CenterCrop):
            pass

        class Compose:
            def __init__(self, ts):
                self.ts = ts

            def __call__(self, image):
                for t in self.ts:
                    image = t(image=image)["image"]
                return {"image": image}

        alb.CenterCrop, alb.RandomCrop, alb.Compose = CenterCrop, RandomCrop, Compose
        sys.modules["albumentations"] = alb

    # the reference's datasets/ has no __init__.py and loses to HuggingFace `datasets`
    ds = types.ModuleType("datasets")
    ds.__path__ = [os.path.join(REF_ROOT, "datasets")]
    sys.modules["datasets"] = ds
    for k in [k for k in sys.modules if k.startswith("datasets.")]:
        del sys.modules[k]

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def import_reference():
    """Returns (vq_module, gpt_module) = the reference's vqvae.big_model_attn_gan and
    transformer.minGPT, imported unmodified."""
    install()
    import importlib

    # make sure we get the reference's top-level packages, not ours
    for name in ("vqvae", "transformer"):
        m = sys.modules.get(name)
        if m is not None and not (getattr(m, "__file__", None) or "").startswith(REF_ROOT) and \
                not any(str(p).startswith(REF_ROOT) for p in getattr(m, "__path__", [])):
            del sys.modules[name]
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        vq = importlib.import_module("vqvae.big_model_attn_gan")
        gpt = importlib.import_module("transformer.minGPT")
    finally:
        os.chdir(cwd)
    return vq, gpt
