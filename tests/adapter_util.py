"""Helpers shared by the adapter tests: golden fixtures of the unmodified reference GaussianAdapter."""
from pathlib import Path

import numpy as np
import torch

from oracle import adapter_ref

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
CASES = ("a", "b")


def load_case(name):
    z = np.load(GOLDEN_DIR / f"adapter_{name}.npz")
    t = lambda k: torch.from_numpy(z[k])
    return dict(z=z, deg=int(z["deg"]), smin=float(z["smin"]), smax=float(z["smax"]),
                image_shape=tuple(int(x) for x in z["image_shape"]),
                extrinsics=t("in_extrinsics"), intrinsics=t("in_intrinsics"), coordinates=t("in_coordinates"),
                depths=t("in_depths"), opacities=t("in_opacities"), raw=t("in_raw"), blocks=t("in_blocks"),
                up={k: t(f"up_{k}") for k in ("means", "covariances", "harmonics")},
                out={k: t(f"out_{k}") for k in ("means", "covariances", "harmonics", "scales", "rotations", "opacities")},
                grad={k: t(f"grad_{k}") for k in ("coordinates", "depths", "raw")})


def oracle_on_case(c, dtype=torch.float64):
    """Runs oracle/adapter_ref.py on a fixture's inputs (flattened to [V,R,..]); returns outputs + autograd grads in
    the fixture's shapes."""
    b, v, r, srf, spp = c["depths"].shape
    V, R = b * v, r * srf
    K = (c["deg"] + 1) ** 2
    coords = c["coordinates"].to(dtype).reshape(V, R, 2).requires_grad_()
    depths = c["depths"].to(dtype).reshape(V, R, spp).requires_grad_()
    raw = c["raw"].to(dtype).reshape(V, R, 7 + 3 * K).requires_grad_()
    out = adapter_ref.adapter_forward(c["extrinsics"].to(dtype).reshape(V, 4, 4), c["intrinsics"].to(dtype).reshape(V, 3, 3),
                                      coords, depths, raw, c["image_shape"], c["deg"], c["smin"], c["smax"],
                                      sh_rotations=c["blocks"].to(dtype))
    shp = (b, v, r, srf, spp)
    res = {k: out[k].reshape(*shp, *out[k].shape[3:]) for k in out}
    loss = sum((res[n] * c["up"][n].to(dtype)).sum() for n in ("means", "covariances", "harmonics"))
    loss.backward()
    grads = dict(coordinates=coords.grad.reshape(c["coordinates"].shape), depths=depths.grad.reshape(c["depths"].shape),
                 raw=raw.grad.reshape(c["raw"].shape))
    return res, grads


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
