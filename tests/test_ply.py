"""`.ply` export / import against the reference's own `export_ply`:
the committed golden vector was produced by the unmodified reference (tools/make_golden_ply.py), and when the
reference tree is present (this container) it is also run live."""
from pathlib import Path

import numpy as np
import pytest
import torch

from ggrt_official_b200 import ply

GOLDEN = Path(__file__).resolve().parent / "golden" / "ply_export.npz"


def _inputs(z):
    return {k: torch.from_numpy(z[k]) for k in ("extrinsics", "means", "scales", "rotations", "harmonics", "opacities")}


def test_vertex_table_matches_the_reference_golden():
    z = np.load(GOLDEN)
    el = ply.vertex_elements(**_inputs(z))
    assert list(el.dtype.names) == list(z["names"]) == ply.construct_list_of_attributes(0)
    table = np.stack([el[n] for n in el.dtype.names], axis=1)
    assert table.dtype == np.float32 and table.shape == z["table"].shape
    np.testing.assert_allclose(table, z["table"], rtol=0, atol=2e-6)


def test_live_reference_export_if_present():
    if not Path("/root/reference/ggrt/model/pixelsplat/ply_export.py").exists():
        pytest.skip("reference tree not present")
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden_ply", Path(__file__).resolve().parent.parent / "tools" / "make_golden_ply.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mod, cap = mg.load_reference_export()
    inp = mg.make_inputs(P=64, K=9, seed=5)
    mod.export_ply(path=Path("/tmp/_unused.ply"), **inp)
    ours = ply.vertex_elements(**inp)
    for n in ours.dtype.names:
        np.testing.assert_allclose(ours[n], cap["elements"][n], rtol=0, atol=2e-6, err_msg=n)


def test_export_import_round_trip(tmp_path):
    z = np.load(GOLDEN)
    inp = _inputs(z)
    path = tmp_path / "sub" / "scene.ply"
    ply.export_ply(path=path, **inp)
    raw = path.read_bytes()
    assert raw.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 257\nproperty float x\n")
    el = ply.read_ply(path)
    written = ply.vertex_elements(**inp)
    np.testing.assert_array_equal(np.stack([el[n] for n in el.dtype.names], 1),
                                  np.stack([written[n] for n in el.dtype.names], 1))
    g = ply.import_ply(path)
    P = inp["means"].shape[0]
    assert g["means"].shape == (P, 3) and g["harmonics"].shape == (P, 3, 1) and g["covariances"].shape == (P, 3, 3)
    np.testing.assert_allclose(g["scales"].log().numpy(), el_cols(el, "scale_"), atol=1e-6)
    # covariances are symmetric PSD with eigenvalues = scales^2
    ev = torch.linalg.eigvalsh(g["covariances"].double())
    np.testing.assert_allclose(np.sort(ev.numpy(), axis=1), np.sort((g["scales"].double() ** 2).numpy(), axis=1),
                               rtol=1e-4, atol=1e-12)
    # imported quaternion is the stored wxyz one, re-ordered to xyzw
    np.testing.assert_array_equal(g["rotations"][:, 3].numpy(), el["rot_0"])


def el_cols(el, prefix):
    return np.stack([el[n] for n in el.dtype.names if n.startswith(prefix)], axis=1)


def test_import_stock_3dgs_layout_with_rest_bands(tmp_path):
    """A stock 3DGS checkpoint carries f_rest_* (channel-major) -- read into harmonics [P,3,K]."""
    P, K = 5, 4
    names = ply.construct_list_of_attributes(3 * (K - 1))
    rng = np.random.default_rng(0)
    el = np.empty(P, dtype=[(n, "<f4") for n in names])
    for n in names:
        el[n] = rng.standard_normal(P).astype(np.float32)
    ply.write_ply(el, tmp_path / "g.ply")
    g = ply.import_ply(tmp_path / "g.ply")
    assert g["harmonics"].shape == (P, 3, K)
    np.testing.assert_array_equal(g["harmonics"][:, 1, 0].numpy(), el["f_dc_1"])
    np.testing.assert_array_equal(g["harmonics"][:, 0, 1].numpy(), el["f_rest_0"])
    np.testing.assert_array_equal(g["harmonics"][:, 1, 1].numpy(), el[f"f_rest_{K - 1}"])
    # ascii variant of the same table
    txt = ["ply", "format ascii 1.0", f"element vertex {P}"] + [f"property float {n}" for n in names] + ["end_header"]
    rows = [" ".join(repr(float(el[n][i])) for n in names) for i in range(P)]
    (tmp_path / "a.ply").write_text("\n".join(txt + rows) + "\n")
    a = ply.read_ply(tmp_path / "a.ply")
    np.testing.assert_allclose(a["opacity"], el["opacity"], rtol=1e-6)
