"""Freezes the CPU oracle's outputs at BASELINE config 1 (10 000 Gaussians, 256x256, 1 view, forward; seed 3407) into
tests/golden/c1_oracle.npz, so that a change of the oracle cannot silently move the target of the GPU parity tests
(SURVEY.md 8c "frozen golden .npz for C1").  Integer results are stored in full, the image as a 64x64 sub-grid plus a
hash of the full buffers."""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ggrt_official_b200.synthetic import SEED, image_gradient, make_scene, to_raster_inputs  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

P, H, W, DEG = 10_000, 256, 256, 4


def compute():
    ri = to_raster_inputs(make_scene(P, H, W, sh_degree=DEG, seed=SEED))
    cam = co.Camera(W=W, H=H, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, view=ri.viewmatrix, proj=ri.projmatrix,
                    campos=ri.campos, bg=ri.bg, deg=DEG)
    f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
    g = image_gradient(H, W, seed=SEED)
    b = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    return dict(
        N=np.int64(f["bin"]["N"]), radii=f["radii"].astype(np.int32), tiles_touched=f["pre"]["tiles_touched"].astype(np.uint32),
        ranges=f["bin"]["ranges"].astype(np.uint32), point_list_sha=sha(f["bin"]["point_list"].astype(np.uint32)),
        keys_sha=sha(f["bin"]["keys"].astype(np.uint64)), n_contrib_sha=sha(f["img"]["n_contrib"].astype(np.uint32)),
        color_grid=f["color"][:, ::4, ::4].astype(np.float32), depth_grid=f["depth"][::4, ::4].astype(np.float32),
        final_T_grid=f["img"]["final_T"][::4, ::4].astype(np.float32),
        color_sum=np.float64(f["color"].astype(np.float64).sum()), fragile=np.int64((f["img"]["fragile"] != 0).sum()),
        dmeans3D_head=b["dmeans3D"][:256].astype(np.float32), dcov3D_head=b["dcov3D"][:256].astype(np.float32),
        dopacity_head=b["dopacity"][:256].astype(np.float32), dsh_head=b["dsh"][:64].astype(np.float32),
        dsh_abs_sum=np.float64(np.abs(b["dsh"].astype(np.float64)).sum()))


if __name__ == "__main__":
    out = ROOT / "tests" / "golden" / "c1_oracle.npz"
    d = compute()
    np.savez_compressed(out, **d)
    print("wrote", out, out.stat().st_size, "bytes; N =", int(d["N"]))
