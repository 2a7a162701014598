"""Distribution of the elementwise gradient error of the CUDA path against the oracle (diagnostics for the bar in
tests/gpu_util.py): for several absolute floors, the worst |got-ref| / (1e-3 |ref| + floor max|ref|) and how many
elements exceed 1, with and without the Gaussians that reach an oracle-flagged fragile pixel."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ggrt_official_b200 import rasterizer as R  # noqa: E402
from ggrt_official_b200.synthetic import image_gradient, make_scene, to_raster_inputs  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from tests import gpu_util as G  # noqa: E402
from tests.helpers import small_case  # noqa: E402


def diag(name, ri, g):
    H, W = ri.image_height, ri.image_width
    st = G.run_cuda_forward(ri)
    cam, f = G.oracle_forward(ri)
    got = R.backward_raw(st, torch.tensor(g, device="cuda:0"))
    torch.cuda.synchronize()
    ref = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs)
    frag = G.fragile_gaussians(f, H, W)
    pairs = dict(dmeans3D=(got["dmeans3D"], ref["dmeans3D"]), dcov3D=(got["dcov3D"], ref["dcov3D"]),
                 dopacity=(got["dopacity"].reshape(-1), ref["dopacity"]), dmeans2D=(got["dmeans2D"][:, :2], ref["dmean2D"]),
                 dsh=(got["dsh"], ref["dsh"]))
    out = {"case": name, "fragile_pixels": int((f["img"]["fragile"] != 0).sum()), "fragile_gaussians": int(frag.sum())}
    for k, (a, b) in pairs.items():
        a = a.detach().cpu().numpy().astype(np.float64).reshape(a.shape[0], -1)
        b = np.asarray(b, np.float64).reshape(a.shape[0], -1)
        scale = np.abs(b).max()
        d = np.abs(a - b)
        row = {"max_rel": float(d.max() / scale)}
        for floor in (1e-6, 1e-5, 1e-4):
            ratio = d / (1e-3 * np.abs(b) + floor * scale)
            row[f"floor{floor:g}"] = {"worst_all": float(ratio.max()), "bad_all": int((ratio > 1).sum()),
                                      "worst_strict": float(ratio[~frag].max()), "bad_strict": int((ratio[~frag] > 1).sum())}
        out[k] = row
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for P, H, W, deg, cs, seed in ((2000, 64, 80, 4, 1.0, 1), (3000, 100, 75, 4, 9.0, 2), (2500, 96, 128, 2, 25.0, 4),
                                   (4000, 32, 32, 0, 60.0, 8)):
        _, ri = small_case(P, H, W, deg, seed=seed, cov_scale=cs)
        diag(f"small P={P} {W}x{H} cs={cs}", ri, np.random.default_rng(seed).standard_normal((3, H, W)).astype(np.float32))
    ri = to_raster_inputs(make_scene(300_000, 756, 1008, sh_degree=4))
    diag("C2", ri, image_gradient(756, 1008))
