"""Host-side API surface of the drop-in package (no GPU)."""
import numpy as np
import pytest
import torch

import diff_gaussian_rasterization as dgr
from ggrt_official_b200 import GaussianRasterizationSettings, GaussianRasterizer
from ggrt_official_b200.view_parallel import GradientArena, shard_views


def _settings(**kw):
    base = dict(image_height=32, image_width=32, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3), scale_modifier=1.0,
                viewmatrix=torch.eye(4), projmatrix=torch.eye(4), sh_degree=0, campos=torch.zeros(3),
                prefiltered=False)
    base.update(kw)
    return GaussianRasterizationSettings(**base)


def test_shim_exports_the_names_ggrt_imports():
    # cuda_splatting.py:6-9
    assert dgr.GaussianRasterizationSettings is GaussianRasterizationSettings
    assert dgr.GaussianRasterizer is GaussianRasterizer


def test_settings_debug_is_optional():
    # render_cuda omits `debug` (cuda_splatting.py:101-113); render_cuda_orthographic passes it (:193-206)
    assert _settings().debug is False
    assert _settings(debug=True).debug is True
    assert _settings()._fields[:11] == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                                        "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered")


def test_exactly_one_of_checks():
    r = GaussianRasterizer(_settings())
    P = 4
    m, o, c6 = torch.zeros(P, 3), torch.zeros(P, 1), torch.zeros(P, 6)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=o, cov3D_precomp=c6)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=o, shs=torch.zeros(P, 1, 3), colors_precomp=torch.zeros(P, 3),
          cov3D_precomp=c6)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m, means2D=m, opacities=o, shs=torch.zeros(P, 1, 3))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m, means2D=m, opacities=o, shs=torch.zeros(P, 1, 3), scales=torch.zeros(P, 3),
          rotations=torch.zeros(P, 4), cov3D_precomp=c6)
    # scales / rotations (the stock 3DGS interface; GGRt never uses it): accepted, the covariance is built with PyTorch
    # and the call then fails only for the same reason every CPU call does
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(means3D=m, means2D=m, opacities=o, shs=torch.zeros(P, 1, 3), scales=torch.zeros(P, 3),
          rotations=torch.zeros(P, 4))
    with pytest.raises(ValueError, match="scales must be"):
        r(means3D=m, means2D=m, opacities=o, shs=torch.zeros(P, 1, 3), scales=torch.zeros(P, 2),
          rotations=torch.zeros(P, 4))


def test_covariance_from_scaling_rotation_is_upstreams_formula():
    """Sigma = R S^2 R^T with R from the quaternion (r, x, y, z), upper triangle in (xx, xy, xz, yy, yz, zz) order."""
    import numpy as np

    from ggrt_official_b200.rasterizer import covariance_from_scaling_rotation

    rng = np.random.default_rng(0)
    q = rng.normal(size=(50, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    s = rng.uniform(0.1, 2.0, size=(50, 3))
    got = covariance_from_scaling_rotation(torch.tensor(s, dtype=torch.float32), torch.tensor(q, dtype=torch.float32), 1.5)
    for i in range(50):
        r_, x, y, z = q[i]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r_ * z), 2 * (x * z + r_ * y)],
                      [2 * (x * y + r_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r_ * x)],
                      [2 * (x * z - r_ * y), 2 * (y * z + r_ * x), 1 - 2 * (x * x + y * y)]])
        assert abs(np.linalg.det(R) - 1) < 1e-9                      # a proper rotation for unit quaternions
        S = np.diag(1.5 * s[i])
        cov = R @ S @ S.T @ R.T
        ref = cov[np.triu_indices(3)]
        assert np.allclose(got[i].numpy(), ref, rtol=1e-5, atol=1e-6)


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly: there is no eager / oracle path inside the product."""
    r = GaussianRasterizer(_settings())
    P = 4
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(means3D=torch.zeros(P, 3), means2D=torch.zeros(P, 3), opacities=torch.zeros(P, 1),
          shs=torch.zeros(P, 1, 3), cov3D_precomp=torch.zeros(P, 6))
    with pytest.raises(RuntimeError, match="no CPU path"):
        r.markVisible(torch.zeros(P, 3))


def test_product_does_not_import_the_oracle():
    import pathlib
    import re

    pkg = pathlib.Path(__file__).resolve().parent.parent / "ggrt_official_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), re.M), f


def test_shard_views_and_arena():
    assert shard_views(8, 0, 8) == [0] and shard_views(8, 7, 8) == [7]
    assert shard_views(5, 1, 2) == [1, 3] and shard_views(1, 1, 2) == []
    assert sorted(sum((shard_views(7, r, 3) for r in range(3)), [])) == list(range(7))
    a = GradientArena.allocate(10, 25, "cpu")
    assert a.flat.numel() == 10 * (3 + 6 + 1 + 75)
    a.flat.zero_()
    a.views["dsh"].fill_(1.0)
    assert float(a.flat.sum()) == 10 * 75 and a.views["dcov3D"].shape == (10, 6)
    assert a.all_reduce() is None  # no process group: a no-op


def test_debug_mode_from_settings_or_environment(monkeypatch):
    from ggrt_official_b200 import rasterizer as R

    rs = R.GaussianRasterizationSettings(1, 1, 1.0, 1.0, None, 1.0, None, None, 0, None, False)
    assert not R._debug_enabled(rs)
    monkeypatch.setenv("GGRT_RASTER_DEBUG", "1")
    assert R._debug_enabled(rs)
    monkeypatch.delenv("GGRT_RASTER_DEBUG")
    assert R._debug_enabled(rs._replace(debug=True))


def test_capacity_estimate_is_a_decaying_high_water_mark():
    """The pair-buffer estimate a shape's next forward is sized from: the largest recent pair count, forgotten slowly --
    a loop alternating between sparse and dense frames must not overflow its speculative buffer on every dense frame."""
    from ggrt_official_b200 import rasterizer as R

    key = ("test", 1, 2, 3)
    R._capacity_cache.pop(key, None)
    R._remember_counts(key, 1000, 40)
    assert R._capacity_cache[key] == (1000, 40)
    R._remember_counts(key, 5000, 90)          # a denser frame raises the mark at once
    assert R._capacity_cache[key] == (5000, 90)
    R._remember_counts(key, 1000, 40)          # a sparse frame lowers it by a few per cent only
    n, m = R._capacity_cache[key]
    assert 0.95 * 5000 <= n < 5000 and 0.95 * 90 <= m <= 90
    for _ in range(400):                        # ... but a long sparse phase forgets the dense frame
        R._remember_counts(key, 1000, 40)
    assert R._capacity_cache[key] == (1000, 40)
    R._capacity_cache.pop(key, None)


def test_decoder_capture_takes_one_scene():
    """DecoderSplattingCUDA.capture freezes the views of ONE scene into a CUDA graph; a batch of scenes is refused
    before anything touches the device."""
    import pytest
    import torch

    from ggrt_official_b200.decoder import DecoderSplattingCUDA, Gaussians

    g = Gaussians(means=torch.zeros(2, 5, 3), covariances=torch.zeros(2, 5, 3, 3), harmonics=torch.zeros(2, 5, 3, 9),
                  opacities=torch.zeros(2, 5))
    E = torch.eye(4).expand(2, 3, 4, 4)
    K = torch.eye(3).expand(2, 3, 3, 3)
    with pytest.raises(ValueError, match="one scene"):
        DecoderSplattingCUDA().capture(g, E, K, torch.ones(2, 3), torch.ones(2, 3) * 10, (16, 16),
                                       grad_color=torch.zeros(3, 3, 16, 16))
