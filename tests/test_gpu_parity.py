"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): tile indices / radii / sort order bit-exact; RGB and depth
within 1e-4 abs (depth relative to its range); gradients within 1e-3 relative.  Pixels the
oracle flags "fragile" (a threshold test within rounding of its boundary, where a different
exp()/fma rounding may legitimately take the other branch) are excluded from the strict bar
and must be rare.
"""
import numpy as np
import pytest
import torch

from ggrt_official_b200 import GaussianRasterizer
from ggrt_official_b200 import rasterizer as R
from ggrt_official_b200.synthetic import image_gradient, make_scene, to_raster_inputs
from oracle import c_oracle as co
from tests import gpu_util as G
from tests.helpers import oracle_camera, small_case

pytestmark = pytest.mark.gpu

CASES = [
    # P, H, W, deg, bg, cov_scale, seed
    (2000, 64, 80, 4, (0.0, 0.0, 0.0), 1.0, 1),
    (3000, 100, 75, 4, (0.2, 0.5, 0.7), 9.0, 2),  # ragged image size, larger splats, background
    (1500, 48, 48, 0, (1.0, 1.0, 1.0), 4.0, 3),  # degree-0 (the depth pass of render_depth_cuda)
    (2500, 96, 128, 2, (0.0, 0.0, 0.0), 25.0, 4),
    (1000, 33, 47, 3, (0.3, 0.3, 0.3), 1.0, 5),
    (500, 16, 16, 1, (0.0, 0.0, 0.0), 100.0, 6),  # single tile, heavy overlap
    (1400, 32, 32, 1, (0.0, 0.0, 0.0), 60.0, 7),  # 513..1024 pairs per tile: 4-warp register sort
    (1900, 32, 32, 0, (0.0, 0.0, 0.0), 60.0, 9),  # 1025..2048 pairs per tile: 8-warp register sort
    (4000, 32, 32, 0, (0.1, 0.1, 0.1), 60.0, 8),  # > 2048 pairs per tile: shared-memory sort, multi-batch render
]


def _check_forward(res, H, W, dense=True):
    for k in ("radii_mismatch", "rect_mismatch", "tiles_mismatch", "xy_bits_mismatch", "conic_bits_mismatch",
              "depth_bits_mismatch", "starts_mismatch"):
        assert res[k] == 0, (k, res)
    assert res["flags_mismatch"] <= 2, res
    assert res["N"][0] == res["N"][1], res
    if res["N"][0]:
        assert res["point_list_mismatch"] == 0 and res["key_depth_mismatch"] == 0 and res["key_idx_mismatch"] == 0, res
    assert res["rgb_max_err"] < 1e-5, res
    assert res["color_max_err"] < 1e-4, res
    assert res["depth_max_relerr"] < 1e-4, res
    assert res["final_T_max_err"] < 1e-5, res
    assert res["n_contrib_mismatch"] == 0, res
    T = ((W + 15) // 16) * ((H + 15) // 16)
    assert res["fragile_pixels"] <= G.fragile_allowance(H * W, res["N"][1], T, dense=dense), res
    assert res["color_max_err_fragile"] < 2e-2, res  # a flipped 1/255 contribution is bounded


@pytest.mark.parametrize("P,H,W,deg,bg,cov_scale,seed", CASES)
def test_forward_matches_oracle(P, H, W, deg, bg, cov_scale, seed):
    _, ri = small_case(P, H, W, deg, bg=bg, seed=seed, cov_scale=cov_scale)
    st = G.run_cuda_forward(ri, debug=True)
    _, f = G.oracle_forward(ri)
    _check_forward(G.compare_forward(st, f), H, W)


@pytest.mark.parametrize("P,H,W,deg,bg,cov_scale,seed", CASES)
def test_backward_matches_oracle(P, H, W, deg, bg, cov_scale, seed):
    _, ri = small_case(P, H, W, deg, bg=bg, seed=seed, cov_scale=cov_scale)
    st = G.run_cuda_forward(ri)
    cam, f = G.oracle_forward(ri)
    g = np.random.default_rng(seed).standard_normal((3, H, W)).astype(np.float32)
    got = R.backward_raw(st, torch.tensor(g, device="cuda:0"))
    torch.cuda.synchronize()
    ref = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs)
    G.assert_grads(G.grad_errors(got, ref, fragile=G.fragile_gaussians(f, H, W)))


def test_colors_precomp_path():
    _, ri = small_case(1200, 64, 64, 0, bg=(0.1, 0.2, 0.3), seed=9, cov_scale=4.0)
    colors = np.abs(ri.shs[:, 0, :]).copy()
    st = G.run_cuda_forward(ri, colors=colors)
    cam, f = G.oracle_forward(ri, colors=colors)
    res = G.compare_forward(st, f)
    assert res["color_max_err"] < 1e-4 and res["point_list_mismatch"] == 0, res
    g = np.random.default_rng(0).standard_normal((3, 64, 64)).astype(np.float32)
    got = R.backward_raw(st, torch.tensor(g, device="cuda:0"))
    ref = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, colors=colors)
    G.assert_grads(G.grad_errors(got, ref, use_sh=False, fragile=G.fragile_gaussians(f, 64, 64)))


def test_aux_channel_forward_and_backward():
    """The optional 4th blended channel (used to fuse GGRt's depth pass into the colour pass)."""
    P, H, W = 2500, 80, 96
    _, ri = small_case(P, H, W, 4, bg=(0.1, 0.3, 0.2), seed=21, cov_scale=6.0)
    dev = "cuda:0"
    rng = np.random.default_rng(4)
    aux = rng.uniform(0.0, 3.0, P).astype(np.float32)
    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    st = R.forward_raw(t(ri.means3D), t(ri.shs), None, t(ri.opacities), t(ri.cov3D), G.settings_from(ri, dev), aux=t(aux))
    cam = oracle_camera(ri)
    f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs, aux=aux)
    ok = f["img"]["fragile"] == 0
    assert np.abs(st["depth"].cpu().numpy() - f["depth"])[ok].max() < 1e-4
    assert np.abs(st["color"].cpu().numpy() - f["color"])[:, ok].max() < 1e-4
    g = rng.standard_normal((3, H, W)).astype(np.float32)
    ga = rng.standard_normal((H, W)).astype(np.float32)
    got = R.backward_raw(st, t(g), grad_aux=t(ga))
    ref = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs, dL_ddepth_img=ga)
    G.assert_grads(G.grad_errors(got, ref, fragile=G.fragile_gaussians(f, H, W)))
    a, b = got["daux"].cpu().numpy().astype(np.float64), ref["daux"].astype(np.float64)
    assert np.isfinite(a).all() and np.abs(a - b).max() < 1e-3 * np.abs(b).max()
    with pytest.raises(RuntimeError, match="aux_precomp"):
        st2 = R.forward_raw(t(ri.means3D), t(ri.shs), None, t(ri.opacities), t(ri.cov3D), G.settings_from(ri, dev))
        R.backward_raw(st2, t(g), grad_aux=t(ga))


def test_camera_gradients_match_autograd_restatement():
    """Opt-in extension (BASELINE config 3): gradients w.r.t. viewmatrix / projmatrix / campos flow when those
    tensors require grad.  Checked against autograd of the dense PyTorch restatement (oracle/torch_ref.py)."""
    from oracle import torch_ref as tr

    P, H, W, deg = 400, 48, 64, 2
    _, ri = small_case(P, H, W, deg, bg=(0.2, 0.1, 0.3), seed=31, cov_scale=9.0, behind_fraction=0.0)
    dev = "cuda:0"
    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    view, proj, campos = t(ri.viewmatrix).requires_grad_(), t(ri.projmatrix).requires_grad_(), t(ri.campos).requires_grad_()
    rs = G.settings_from(ri, dev)._replace(viewmatrix=view, projmatrix=proj, campos=campos)
    means = t(ri.means3D).requires_grad_()
    img, _, _ = GaussianRasterizer(rs)(means3D=means, means2D=torch.zeros_like(means, requires_grad=True),
                                       shs=t(ri.shs), opacities=t(ri.opacities), cov3D_precomp=t(ri.cov3D))
    g = np.random.default_rng(8).standard_normal((3, H, W)).astype(np.float32)
    (img * t(g)).sum().backward()

    d = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    v64, p64, c64, m64 = d(ri.viewmatrix).requires_grad_(), d(ri.projmatrix).requires_grad_(), \
        d(ri.campos).requires_grad_(), d(ri.means3D).requires_grad_()
    color, _, _ = tr.rasterize(m64, d(ri.cov3D), d(ri.opacities), view=v64, proj=p64, campos=c64, bg=d(ri.bg),
                               tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, H=H, W=W, deg=deg, sh=d(ri.shs))
    (color * d(g)).sum().backward()
    for name, got, ref in (("viewmatrix", view.grad, v64.grad), ("projmatrix", proj.grad, p64.grad),
                           ("campos", campos.grad, c64.grad), ("means3D", means.grad, m64.grad)):
        got, ref = got.cpu().double(), ref
        err = float((got - ref).abs().max() / ref.abs().max())
        assert err < 2e-3, (name, err, got, ref)
    # camera tensors that do not require grad cost nothing and get no gradient
    rs2 = G.settings_from(ri, dev)
    img2, _, _ = GaussianRasterizer(rs2)(means3D=means, means2D=torch.zeros_like(means, requires_grad=True),
                                         shs=t(ri.shs), opacities=t(ri.opacities), cov3D_precomp=t(ri.cov3D))
    img2.sum().backward()
    assert rs2.viewmatrix.grad is None


def test_speculative_binning_overflow_is_redone():
    """After the first call of a shape the pair buffer is sized from the previous N (no host wait).  If the next
    scene of the same shape needs more pairs than guessed, the binning + render is redone: results stay exact."""
    P, H, W = 3000, 96, 96
    _, sparse = small_case(P, H, W, 2, seed=41, cov_scale=1.0)
    _, dense = small_case(P, H, W, 2, seed=42, cov_scale=64.0)
    R._capacity_cache.clear()
    st0 = G.run_cuda_forward(sparse)        # first call of this shape: exact, synchronous
    st1 = G.run_cuda_forward(dense)         # guess from the sparse scene is far too small -> redo path
    assert st1["N"] > 2 * st0["N"] and st1["capacity"] == st1["N"]
    _, f1 = G.oracle_forward(dense)
    _check_forward(G.compare_forward(st1, f1), H, W)
    st2 = G.run_cuda_forward(dense)         # now the guess is large enough: speculative path, capacity > N
    assert st2["capacity"] > st2["N"] == st1["N"]
    _check_forward(G.compare_forward(st2, f1), H, W)
    g = np.random.default_rng(1).standard_normal((3, H, W)).astype(np.float32)
    cam = oracle_camera(dense)
    ref = co.backward(cam, dense.means3D, dense.cov3D, dense.opacities, f1, g, sh=dense.shs)
    for st in (st1, st2):
        got = R.backward_raw(st, torch.tensor(g, device="cuda:0"))
        G.assert_grads(G.grad_errors(got, ref, fragile=G.fragile_gaussians(f1, H, W)))
    st3 = G.run_cuda_forward(sparse)        # shrinking scenes are fine too
    _, f0 = G.oracle_forward(sparse)
    _check_forward(G.compare_forward(st3, f0), H, W)


def test_config1_10k_256(tmp_path):
    """BASELINE config 1: 10K Gaussians, 256x256, 1 view, forward vs the oracle."""
    ri = to_raster_inputs(make_scene(10_000, 256, 256, sh_degree=4))
    st = G.run_cuda_forward(ri)
    _, f = G.oracle_forward(ri)
    _check_forward(G.compare_forward(st, f), 256, 256, dense=False)


def test_autograd_module_end_to_end():
    """Through GaussianRasterizer / autograd exactly as cuda_splatting.py:101-125 calls it."""
    P, H, W = 3000, 72, 96
    _, ri = small_case(P, H, W, 4, bg=(0.05, 0.1, 0.15), seed=11, cov_scale=4.0)
    dev = "cuda:0"
    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    means, cov, opac, shs = (t(ri.means3D).requires_grad_(), t(ri.cov3D).requires_grad_(),
                             t(ri.opacities).requires_grad_(), t(ri.shs).requires_grad_())
    means2D = torch.zeros_like(means, requires_grad=True)
    rs = G.settings_from(ri, dev)
    # campos as the stride-4 column slice the reference passes (cuda_splatting.py:111)
    c2w = torch.eye(4, device=dev)
    c2w[:3, 3] = t(ri.campos)
    rs = rs._replace(campos=c2w[:3, 3])
    assert not rs.campos.is_contiguous()
    image, radii, depth = GaussianRasterizer(rs)(means3D=means, means2D=means2D, shs=shs, colors_precomp=None,
                                                  opacities=opac, cov3D_precomp=cov)
    assert image.shape == (3, H, W) and radii.shape == (P,) and radii.dtype == torch.int32 and depth.shape == (H, W)
    g = np.random.default_rng(3).standard_normal((3, H, W)).astype(np.float32)
    (image * t(g)).sum().backward()
    cam = oracle_camera(ri)
    f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
    ref = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs)
    ok = f["img"]["fragile"] == 0
    assert np.abs(image.detach().cpu().numpy() - f["color"])[:, ok].max() < 1e-4
    got = dict(dmeans3D=means.grad, dcov3D=cov.grad, dopacity=opac.grad, dmeans2D=means2D.grad, dsh=shs.grad)
    G.assert_grads(G.grad_errors(got, ref, fragile=G.fragile_gaussians(f, H, W)))


def test_empty_and_culled_inputs():
    dev = "cuda:0"
    _, ri = small_case(64, 32, 32, 0, seed=1, behind_fraction=0.0)
    rs = G.settings_from(ri, dev)
    # every Gaussian behind the camera (point reflection through the camera centre): image is the
    # background, radii all zero, gradients zero
    means = (2 * torch.tensor(ri.campos, device=dev)[None] - torch.tensor(ri.means3D, device=dev)).requires_grad_()
    rs = rs._replace(bg=torch.tensor([0.25, 0.5, 0.75], device=dev))
    img, radii, depth = GaussianRasterizer(rs)(means3D=means, means2D=torch.zeros_like(means, requires_grad=True),
                                                shs=torch.tensor(ri.shs, device=dev), opacities=torch.tensor(ri.opacities, device=dev),
                                                cov3D_precomp=torch.tensor(ri.cov3D, device=dev))
    assert int(radii.abs().sum()) == 0
    assert torch.allclose(img, torch.tensor([0.25, 0.5, 0.75], device=dev)[:, None, None].expand_as(img))
    assert float(depth.abs().max()) == 0.0
    img.sum().backward()
    assert float(means.grad.abs().max()) == 0.0
    # P = 0
    z = lambda *s: torch.zeros(*s, device=dev)
    img0, radii0, _ = GaussianRasterizer(rs)(means3D=z(0, 3), means2D=z(0, 3), shs=z(0, 1, 3), opacities=z(0, 1),
                                             cov3D_precomp=z(0, 6))
    assert radii0.numel() == 0 and torch.allclose(img0[:, 0, 0], rs.bg)


def test_mark_visible():
    dev = "cuda:0"
    _, ri = small_case(5000, 64, 64, 0, seed=2, behind_fraction=0.3)
    rs = G.settings_from(ri, dev)
    vis = GaussianRasterizer(rs).markVisible(torch.tensor(ri.means3D, device=dev)).cpu().numpy()
    assert np.array_equal(vis, co.mark_visible(oracle_camera(ri), ri.means3D))
