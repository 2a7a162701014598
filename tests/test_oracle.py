"""The C oracle against (i) the independent autograd restatement, (ii) closed-form cases.

PARITY UNPINNED: the reference ships no golden vectors for the rasterizer;
these tests pin the oracle to itself through two independent restatements and analytic facts.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import torch_ref as tr
from tests.helpers import oracle_camera, rel_err, small_case


def _t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64)


@pytest.mark.parametrize(
    "P,H,W,deg,bg,use_sh,seed",
    [
        (150, 40, 56, 4, (0.2, 0.5, 0.7), True, 1),
        (300, 64, 64, 3, (0.0, 0.0, 0.0), True, 2),
        (80, 33, 47, 0, (1.0, 1.0, 1.0), True, 3),
        (100, 48, 48, 2, (0.1, 0.2, 0.3), False, 4),
        (120, 32, 32, 1, (0.0, 0.0, 0.0), True, 5),
    ],
)
def test_c_oracle_matches_autograd_restatement(P, H, W, deg, bg, use_sh, seed):
    _, ri = small_case(P, H, W, deg, bg=bg, seed=seed, cov_scale=9.0)
    cam = oracle_camera(ri)
    colors = None if use_sh else np.abs(ri.shs[:, 0, :]).copy()
    sh = ri.shs if use_sh else None
    f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=sh, colors=colors)
    rng = np.random.default_rng(5)
    g = rng.standard_normal((3, H, W)).astype(np.float32)
    gd = rng.standard_normal((H, W)).astype(np.float32)
    b = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=sh, colors=colors, dL_ddepth_img=gd)

    means, cv, op = _t(ri.means3D).requires_grad_(), _t(ri.cov3D).requires_grad_(), _t(ri.opacities).requires_grad_()
    sht = _t(ri.shs).requires_grad_() if use_sh else None
    col = _t(colors).requires_grad_() if not use_sh else None
    color, depth, aux = tr.rasterize(means, cv, op, view=_t(ri.viewmatrix), proj=_t(ri.projmatrix),
                                     campos=_t(ri.campos), bg=_t(ri.bg), tanfovx=ri.tanfovx, tanfovy=ri.tanfovy,
                                     H=H, W=W, deg=deg, sh=sht, colors=col)
    ((color * _t(g)).sum() + (depth * _t(gd)).sum()).backward()

    assert f["bin"]["N"] > 2 * P * 0.5
    vis = f["radii"] > 0
    assert np.array_equal(aux["radii"].numpy(), f["radii"])
    assert np.array_equal(aux["rect"].numpy()[vis], f["pre"]["rect"][vis])
    ok = f["img"]["fragile"] == 0
    assert ok.mean() > 0.99
    assert np.abs(color.detach().numpy() - f["color"])[:, ok].max() < 2e-6
    assert np.abs(depth.detach().numpy() - f["depth"])[ok].max() < 5e-5
    assert np.allclose(aux["final_T"].detach().numpy()[ok], f["img"]["final_T"][ok], atol=2e-6)
    # hand-derived backward (A.4/A.5) vs autograd of the forward
    assert rel_err(b["dmeans3D"], means.grad.numpy()) < 2e-5
    assert rel_err(b["dcov3D"], cv.grad.numpy()) < 2e-5
    assert rel_err(b["dopacity"], op.grad.numpy().reshape(-1)) < 2e-5
    if use_sh:
        assert rel_err(b["dsh"], sht.grad.numpy()) < 2e-5
    else:
        assert rel_err(b["dcolor"], col.grad.numpy()) < 2e-5


@pytest.mark.parametrize("P,H,W,deg,seed", [(200, 40, 56, 2, 11), (300, 48, 64, 4, 12), (120, 32, 32, 0, 13)])
def test_c_oracle_camera_gradients_match_autograd(P, H, W, deg, seed):
    """The hand-derived dL/d(viewmatrix, projmatrix, campos) of the C oracle (pose gradients of BASELINE
    config 3) against autograd of the dense float64 restatement.  Pins the target of the GPU pose tests."""
    _, ri = small_case(P, H, W, deg, bg=(0.2, 0.1, 0.3), seed=seed, cov_scale=9.0, behind_fraction=0.0)
    cam = oracle_camera(ri)
    f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
    g = np.random.default_rng(seed).standard_normal((3, H, W)).astype(np.float32)
    b = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs, want_camera=True)
    v64, p64, c64 = _t(ri.viewmatrix).requires_grad_(), _t(ri.projmatrix).requires_grad_(), _t(ri.campos).requires_grad_()
    color, _, _ = tr.rasterize(_t(ri.means3D), _t(ri.cov3D), _t(ri.opacities), view=v64, proj=p64, campos=c64,
                               bg=_t(ri.bg), tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, H=H, W=W, deg=deg, sh=_t(ri.shs))
    (color * _t(g)).sum().backward()
    cam_g = b["dcamera"]
    assert rel_err(cam_g[:16].reshape(4, 4), v64.grad.numpy()) < 1e-4
    assert rel_err(cam_g[16:32].reshape(4, 4), p64.grad.numpy()) < 1e-4
    if deg > 0:
        assert rel_err(cam_g[32:35], c64.grad.numpy()) < 1e-4
    else:  # degree 0 has no view-direction dependence
        assert c64.grad is None and np.abs(cam_g[32:35]).max() == 0.0
    # without the flag nothing is computed
    assert co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs)["dcamera"] is None


def test_sh_basis_is_orthonormal():
    """Degree 0..4 real SH basis (A.1 constants) integrates to the identity on the sphere."""
    n = 64
    x, w = np.polynomial.legendre.leggauss(n)
    phi = (np.arange(2 * n) + 0.5) * np.pi / n
    ct, ph = np.meshgrid(x, phi, indexing="ij")
    st = np.sqrt(1 - ct * ct)
    d = np.stack([st * np.cos(ph), st * np.sin(ph), ct], -1).reshape(-1, 3)
    wt = np.repeat(w, 2 * n) * (np.pi / n)
    B = tr.sh_basis(4, torch.tensor(d)).numpy()
    G = (B * wt[:, None]).T @ B
    assert np.abs(G - np.eye(25)).max() < 1e-12


def _single(mean_cam, cov6, opacity, H=32, W=32, sh0=(1.0, 2.0, 3.0), bg=(0.0, 0.0, 0.0)):
    view = np.eye(4, dtype=np.float32)
    f = 20.0  # focal in pixels
    tanx, tany = W / (2 * f), H / (2 * f)
    proj = np.zeros((4, 4), np.float32)  # p_hom = [p,1] . proj, ndc = x / (z tan)
    proj[0, 0], proj[1, 1], proj[2, 2], proj[2, 3], proj[3, 2] = 1 / tanx, 1 / tany, 1.0, 1.0, -0.01
    cam = co.Camera(W=W, H=H, tanfovx=tanx, tanfovy=tany, view=view, proj=proj, campos=np.zeros(3, np.float32),
                    bg=np.asarray(bg, np.float32), deg=0)
    sh = (np.asarray(sh0, np.float32)[None, None, :] - 0.5) / 0.28209479177387814
    return cam, co.forward(cam, np.asarray([mean_cam], np.float32), np.asarray([cov6], np.float32),
                           np.asarray([opacity], np.float32), sh=sh.astype(np.float32))


def test_single_gaussian_closed_form():
    """One isotropic Gaussian on the optical axis: alpha(x) = o exp(-r^2 / 2 s^2), s^2 = (f sigma/z)^2 + 0.3."""
    z, sig, o, f = 4.0, 0.5, 0.8, 20.0
    cam, out = _single((0.0, 0.0, z), (sig**2, 0, 0, sig**2, 0, sig**2), o)
    s2 = (f * sig / z) ** 2 + 0.3
    assert out["radii"][0] == int(np.ceil(3 * np.sqrt(s2)))
    cx, cy = (cam.W - 1) / 2, (cam.H - 1) / 2
    assert np.allclose(out["pre"]["xy"][0], [cx, cy], atol=1e-4)
    ys, xs = np.mgrid[: cam.H, : cam.W]
    alpha = o * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * s2))
    alpha = np.where(alpha >= 1 / 255, np.minimum(alpha, 0.99), 0.0)
    ok = out["img"]["fragile"] == 0
    for ch, c in enumerate((1.0, 2.0, 3.0)):
        assert np.abs(out["color"][ch] - c * alpha)[ok].max() < 1e-5
    assert np.abs(out["depth"] - z * alpha)[ok].max() < 1e-5
    assert np.abs(out["img"]["final_T"] - (1 - alpha))[ok].max() < 1e-6


def test_cull_cases():
    s = 0.01
    iso = (s, 0, 0, s, 0, s)
    assert _single((0, 0, 0.2), iso, 0.5)[1]["radii"][0] == 0  # z <= 0.2 culled (A.1)
    assert _single((0, 0, -3.0), iso, 0.5)[1]["radii"][0] == 0  # behind camera
    assert _single((0, 0, 0.2001), iso, 0.5)[1]["radii"][0] > 0
    assert _single((100.0, 0, 4.0), iso, 0.5)[1]["radii"][0] == 0  # far off-screen: empty tile rect
    cam, out = _single((0, 0, 4.0), iso, 0.5, bg=(0.25, 0.5, 0.75))
    assert np.allclose(out["color"][:, 0, 0], [0.25, 0.5, 0.75])  # untouched corner shows background


def test_equal_depth_tie_order_and_ranges():
    """Two Gaussians at identical depth keep ascending index order inside each tile list (A.2)."""
    view = np.eye(4, dtype=np.float32)
    W = H = 32
    f = 20.0
    tan = W / (2 * f)
    proj = np.zeros((4, 4), np.float32)
    proj[0, 0] = proj[1, 1] = 1 / tan
    proj[2, 2] = proj[2, 3] = 1.0
    cam = co.Camera(W=W, H=H, tanfovx=tan, tanfovy=tan, view=view, proj=proj, campos=np.zeros(3, np.float32), deg=0)
    means = np.asarray([[0.1, 0, 4], [0.0, 0, 4], [0.0, 0.1, 2]], np.float32)
    cov = np.tile(np.asarray([[0.04, 0, 0, 0.04, 0, 0.04]], np.float32), (3, 1))
    out = co.forward(cam, means, cov, np.full(3, 0.5, np.float32), colors=np.eye(3, dtype=np.float32))
    b = out["bin"]
    assert b["N"] == int(out["pre"]["tiles_touched"].sum())
    for t in range(4):
        lo, hi = b["ranges"][t]
        lst = b["point_list"][lo:hi].tolist()
        assert lst == [2, 0, 1], lst  # nearest first, then the depth tie in index order
    assert np.all(np.diff(b["keys"].astype(np.uint64)) >= 0)


def test_mark_visible():
    _, ri = small_case(500, 64, 64, 0, behind_fraction=0.3)
    cam = oracle_camera(ri)
    vis = co.mark_visible(cam, ri.means3D)
    tz = ri.means3D @ ri.viewmatrix[:3, 2] + ri.viewmatrix[3, 2]
    assert 0.5 < vis.mean() < 0.9
    assert np.array_equal(vis[np.abs(tz - 0.2) > 1e-4], (tz > 0.2)[np.abs(tz - 0.2) > 1e-4])


def test_oracle_reproduces_its_frozen_c1_golden():
    """BASELINE config 1 (10K Gaussians, 256x256): the oracle against its own frozen outputs -- guards the target of
    the GPU parity tests against silent drift (tools/make_golden_c1.py)."""
    import importlib.util
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("make_golden_c1", root / "tools" / "make_golden_c1.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    got = mg.compute()
    z = np.load(root / "tests" / "golden" / "c1_oracle.npz")
    for k in ("N", "radii", "tiles_touched", "ranges", "fragile"):
        assert np.array_equal(got[k], z[k]), k
    for k in ("point_list_sha", "keys_sha", "n_contrib_sha"):
        assert str(got[k]) == str(z[k]), k
    for k in ("color_grid", "depth_grid", "final_T_grid"):
        np.testing.assert_allclose(got[k], z[k], rtol=0, atol=1e-6, err_msg=k)
    assert abs(float(got["color_sum"]) - float(z["color_sum"])) <= 1e-6 * abs(float(z["color_sum"]))
    for k in ("dmeans3D_head", "dcov3D_head", "dopacity_head", "dsh_head"):
        scale = max(float(np.abs(z[k]).max()), 1e-30)
        assert float(np.abs(got[k] - z[k]).max()) <= 1e-5 * scale, k
    assert abs(float(got["dsh_abs_sum"]) - float(z["dsh_abs_sum"])) <= 1e-5 * float(z["dsh_abs_sum"])
