"""Imports the UNMODIFIED reference render glue from /root/reference (container only).

The reference package pulls in heavy optional dependencies at import time (e3nn, omegaconf,
matplotlib, ...) that are unrelated to the render path and absent here; they are stubbed in
sys.modules (SURVEY.md 8b "Empirical boundary probe").  Returns None when the reference
tree is not present (e.g. on the GPU box)."""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REF = Path("/root/reference")
STUBS = ("e3nn", "e3nn.o3", "omegaconf", "colorspacious", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
         "plyfile", "lpips", "imageio", "visdom", "tensorboardX", "hydra", "moviepy", "moviepy.editor", "dacite",
         "scipy.spatial.transform", "skimage", "skimage.metrics", "skimage.io", "cv2", "tabulate", "wandb",
         "pytorch_lightning", "lightning_fabric", "kornia")


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


def load_reference_glue():
    """-> (cuda_splatting module, decoder_splatting_cuda module) of the reference, or None."""
    if not (REF / "ggrt" / "model" / "pixelsplat" / "decoder" / "cuda_splatting.py").exists():
        return None
    root = str(Path(__file__).resolve().parent.parent)
    if root not in sys.path:
        sys.path.insert(0, root)  # provides the diff_gaussian_rasterization shim
    if str(REF) not in sys.path:
        sys.path.append(str(REF))
    for name in STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    cs = importlib.import_module("ggrt.model.pixelsplat.decoder.cuda_splatting")
    dec = importlib.import_module("ggrt.model.pixelsplat.decoder.decoder_splatting_cuda")
    return cs, dec
