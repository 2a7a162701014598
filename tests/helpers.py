"""Shared helpers for the parity tests (oracle side)."""
from __future__ import annotations

import numpy as np

from ggrt_official_b200.synthetic import make_scene, to_raster_inputs
from oracle import c_oracle as co


def oracle_camera(ri) -> co.Camera:
    return co.Camera(W=ri.image_width, H=ri.image_height, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, view=ri.viewmatrix,
                     proj=ri.projmatrix, campos=ri.campos, bg=ri.bg, deg=ri.sh_degree)


def small_case(P, H, W, deg, bg=(0.0, 0.0, 0.0), seed=1, cov_scale=1.0, behind_fraction=0.05, opacity_max=1 / 3):
    sc = make_scene(P, H, W, sh_degree=deg, seed=seed, behind_fraction=behind_fraction, opacity_max=opacity_max)
    ri = to_raster_inputs(sc, bg=bg)
    ri.cov3D = np.ascontiguousarray(ri.cov3D * np.float32(cov_scale))
    return sc, ri


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
