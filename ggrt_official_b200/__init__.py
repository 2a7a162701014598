"""B200-native differentiable 3D-Gaussian tile rasterizer: a drop-in for the
`diff_gaussian_rasterization` package on GGRt's render path."""
from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
__version__ = "0.1.0"
