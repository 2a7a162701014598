"""Import shim: makes the B200 rasterizer importable under the name GGRt uses
(/root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py:6-9), so the reference
caller runs unmodified with this repository on PYTHONPATH."""
from ggrt_official_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
