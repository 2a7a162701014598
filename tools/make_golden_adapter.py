"""Golden vectors for the fused Gaussian adapter: runs the UNMODIFIED reference `GaussianAdapter`
(/root/reference/ggrt/model/pixelsplat/encoder/common/gaussian_adapter.py) with the call shapes of
encoder_epipolar.py:221-229 and stores inputs, outputs and autograd gradients in tests/golden/adapter_*.npz.

e3nn is not installed here, so the two e3nn functions `rotate_sh` calls are stood in for: `matrix_to_angles`
returns the rotation matrices themselves and `wigner_D(l, ...)` looks the (2l+1)x(2l+1) block up in a table of
fixed random per-view matrices that is stored in the fixture.  The reference's own `rotate_sh` (per-degree einsum
and concatenation, sh_rotation.py:10-29), `quaternion_to_matrix`, `build_covariance`, `get_world_rays` etc. run as
they are.  Container only."""
import importlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
BLOCKS = {}  # set per case: [V,K,K]


def load_reference_adapter():
    o3 = types.ModuleType("e3nn.o3")
    o3.matrix_to_angles = lambda rot: (rot, rot, rot)

    def wigner_D(degree, alpha, beta, gamma):
        l = int(degree)
        full = BLOCKS["full"]  # [b*v,K,K]; alpha is the rotation tensor [b,v,1,1,1,1,3,3]
        lead = alpha.shape[:-2]
        blk = full[:, l * l:(l + 1) ** 2, l * l:(l + 1) ** 2]
        return blk.reshape(*lead, 2 * l + 1, 2 * l + 1)  # lead = (b, v, 1, 1, 1, 1) as rotate_sh is called (:90)

    o3.wigner_D = wigner_D
    e3 = types.ModuleType("e3nn")
    e3.o3 = o3
    sys.modules["e3nn"], sys.modules["e3nn.o3"] = e3, o3
    if str(REF) not in sys.path:
        sys.path.append(str(REF))
    # import the three leaf modules directly (the package __init__ files pull in the whole model zoo)
    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, REF / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    for pkg in ("ggrt", "ggrt.geometry", "ggrt.misc", "ggrt.model", "ggrt.model.pixelsplat", "ggrt.model.pixelsplat.encoder",
                "ggrt.model.pixelsplat.encoder.common"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    load("ggrt.geometry.projection", "ggrt/geometry/projection.py")
    load("ggrt.misc.sh_rotation", "ggrt/misc/sh_rotation.py")
    load("ggrt.model.pixelsplat.encoder.common.gaussians", "ggrt/model/pixelsplat/encoder/common/gaussians.py")
    return load("ggrt.model.pixelsplat.encoder.common.gaussian_adapter",
                "ggrt/model/pixelsplat/encoder/common/gaussian_adapter.py")


def make_case(seed, b, v, h, w, srf, spp, deg, smin=0.5, smax=15.0):
    g = torch.Generator().manual_seed(seed)
    K = (deg + 1) ** 2
    r = h * w
    extr = torch.eye(4).repeat(b, v, 1, 1)
    for i in range(b):
        for j in range(v):
            q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
            if torch.det(q) < 0:
                q[:, 0] = -q[:, 0]
            extr[i, j, :3, :3] = q
            extr[i, j, :3, 3] = torch.randn(3, generator=g)
    intr = torch.eye(3).repeat(b, v, 1, 1)
    intr[..., 0, 0] = 0.8 + 0.2 * torch.rand(b, v, generator=g)
    intr[..., 1, 1] = 1.0 + 0.2 * torch.rand(b, v, generator=g)
    intr[..., 0, 2] = 0.5 + 0.02 * torch.randn(b, v, generator=g)
    intr[..., 1, 2] = 0.5 + 0.02 * torch.randn(b, v, generator=g)
    coords = torch.rand(b, v, r, srf, 1, 2, generator=g)
    depths = 1.0 + 4.0 * torch.rand(b, v, r, srf, spp, generator=g)
    opac = torch.rand(b, v, r, srf, spp, generator=g)
    raw = torch.randn(b, v, r, srf, 1, 7 + 3 * K, generator=g)
    blocks = torch.zeros(b * v, K, K)
    for l in range(deg + 1):
        blocks[:, l * l:(l + 1) ** 2, l * l:(l + 1) ** 2] = torch.randn(b * v, 2 * l + 1, 2 * l + 1, generator=g)
    up = dict(means=torch.randn(b, v, r, srf, spp, 3, generator=g), covariances=torch.randn(b, v, r, srf, spp, 3, 3, generator=g),
              harmonics=torch.randn(b, v, r, srf, spp, 3, K, generator=g))
    return dict(extrinsics=extr, intrinsics=intr, coordinates=coords, depths=depths, opacities=opac, raw=raw,
                blocks=blocks, up=up, image_shape=(h, w), deg=deg, smin=smin, smax=smax)


def run_reference(mod, case):
    BLOCKS["full"] = case["blocks"]
    cfg = mod.GaussianAdapterCfg(gaussian_scale_min=case["smin"], gaussian_scale_max=case["smax"], sh_degree=case["deg"])
    adapter = mod.GaussianAdapter(cfg)
    coords = case["coordinates"].clone().requires_grad_()
    depths = case["depths"].clone().requires_grad_()
    raw = case["raw"].clone().requires_grad_()
    e = case["extrinsics"][:, :, None, None, None]
    k = case["intrinsics"][:, :, None, None, None]
    out = adapter.forward(e, k, coords, depths, case["opacities"], raw, case["image_shape"])
    loss = sum((getattr(out, n) * case["up"][n]).sum() for n in ("means", "covariances", "harmonics"))
    loss.backward()
    return out, dict(coordinates=coords.grad, depths=depths.grad, raw=raw.grad)


CASES = {"a": dict(seed=3407, b=1, v=2, h=6, w=8, srf=1, spp=3, deg=4),
         "b": dict(seed=11, b=2, v=1, h=5, w=7, srf=2, spp=1, deg=2)}

if __name__ == "__main__":
    mod = load_reference_adapter()
    for name, kw in CASES.items():
        case = make_case(**kw)
        out, grads = run_reference(mod, case)
        path = ROOT / "tests" / "golden" / f"adapter_{name}.npz"
        np.savez_compressed(
            path, image_shape=np.array(case["image_shape"]), deg=case["deg"], smin=case["smin"], smax=case["smax"],
            **{f"in_{k}": case[k].numpy() for k in ("extrinsics", "intrinsics", "coordinates", "depths", "opacities", "raw", "blocks")},
            **{f"up_{k}": v.numpy() for k, v in case["up"].items()},
            **{f"out_{k}": getattr(out, k).detach().numpy() for k in ("means", "covariances", "harmonics", "scales", "rotations", "opacities")},
            **{f"grad_{k}": v.numpy() for k, v in grads.items()})
        print("wrote", path, {k: tuple(getattr(out, k).shape) for k in ("means", "covariances", "harmonics", "scales", "rotations")})
