"""Golden vector for ggrt_official_b200.ply: runs the UNMODIFIED reference `export_ply`
(/root/reference/ggrt/model/pixelsplat/ply_export.py:26-92) with a recording stand-in for the absent `plyfile`
package and stores its inputs and the vertex array it hands to PlyElement.describe.  Container only (the
reference tree is not on the GPU box); output: tests/golden/ply_export.npz."""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/ggrt/model/pixelsplat/ply_export.py")


def load_reference_export():
    captured = {}
    ply = types.ModuleType("plyfile")

    class PlyElement:
        @staticmethod
        def describe(elements, name):
            captured["elements"], captured["name"] = elements.copy(), name
            return (elements, name)

    class PlyData:
        def __init__(self, els, **kw):
            self.els = els

        def write(self, path):
            captured["path"] = str(path)

    ply.PlyElement, ply.PlyData = PlyElement, PlyData
    sys.modules["plyfile"] = ply
    if "jaxtyping" not in sys.modules:
        try:
            import jaxtyping  # noqa: F401
        except Exception:
            jt = types.ModuleType("jaxtyping")

            class _F:
                def __class_getitem__(cls, item):
                    return cls

            jt.Float = _F
            sys.modules["jaxtyping"] = jt
    spec = importlib.util.spec_from_file_location("ref_ply_export", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, captured


def make_inputs(P=257, K=25, seed=3407):
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    extr = torch.eye(4)
    extr[:3, :3] = q
    extr[:3, 3] = torch.randn(3, generator=g)
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=-1, keepdim=True)
    return dict(extrinsics=extr, means=torch.randn(P, 3, generator=g) * 3 + 1, scales=torch.rand(P, 3, generator=g) * 0.1 + 1e-3,
                rotations=rot, harmonics=torch.randn(P, 3, K, generator=g), opacities=torch.rand(P, generator=g))


if __name__ == "__main__":
    mod, cap = load_reference_export()
    inp = make_inputs()
    mod.export_ply(path=Path("/tmp/ref_golden.ply"), **inp)
    el = cap["elements"]
    assert cap["name"] == "vertex"
    out = ROOT / "tests" / "golden" / "ply_export.npz"
    np.savez_compressed(out, names=np.array(el.dtype.names), table=np.stack([el[n] for n in el.dtype.names], axis=1),
                        **{k: v.numpy() for k, v in inp.items()})
    print("wrote", out, el.shape, el.dtype.names)
