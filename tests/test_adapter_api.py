"""Host-side behaviour of the fused adapter's Python mirror (no GPU): API surface of the reference class
(gaussian_adapter.py:30-118), loud failure without CUDA, optional e3nn dependency."""
import importlib.util

import pytest
import torch

from ggrt_official_b200.adapter import (GaussianAdapter, GaussianAdapterCfg, Gaussians, sh_rotation_matrices_e3nn)
from tests.adapter_util import load_case


def test_api_surface_matches_the_reference_class():
    ad = GaussianAdapter(GaussianAdapterCfg(gaussian_scale_min=0.5, gaussian_scale_max=15.0, sh_degree=4))
    assert ad.d_sh == 25 and ad.d_in == 7 + 75
    mask = ad.sh_mask
    assert mask.shape == (25,) and float(mask[0]) == 1.0
    for deg in range(1, 5):
        assert torch.allclose(mask[deg ** 2:(deg + 1) ** 2], torch.tensor(0.1 * 0.25 ** deg))
    assert "sh_mask" not in ad.state_dict()  # persistent=False, as in the reference
    assert [f.name for f in Gaussians.__dataclass_fields__.values()] == [
        "means", "covariances", "scales", "rotations", "harmonics", "opacities"]
    with pytest.raises(ValueError):
        GaussianAdapter(GaussianAdapterCfg(0.5, 15.0, 5))


def test_scale_multiplier_matches_the_golden_scales():
    """scales = (min + (max-min) sigmoid(raw)) * depth * multiplier: recover the multiplier from the reference output."""
    c = load_case("a")
    ad = GaussianAdapter(GaussianAdapterCfg(c["smin"], c["smax"], c["deg"]))
    h, w = c["image_shape"]
    mult = ad.get_scale_multiplier(c["intrinsics"], 1 / torch.tensor((w, h), dtype=torch.float32))  # [b, v]
    base = c["smin"] + (c["smax"] - c["smin"]) * c["raw"][..., :3].sigmoid()                          # [b,v,r,srf,1,3]
    expect = base * c["depths"][..., None] * mult[:, :, None, None, None, None]
    assert torch.allclose(expect, c["out"]["scales"], rtol=1e-5, atol=1e-8)


def test_no_cpu_path_and_optional_e3nn():
    c = load_case("b")
    ad = GaussianAdapter(GaussianAdapterCfg(c["smin"], c["smax"], c["deg"]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ad(c["extrinsics"][:, :, None, None, None], c["intrinsics"][:, :, None, None, None], c["coordinates"], c["depths"],
           c["opacities"], c["raw"], c["image_shape"])
    import sys

    if "e3nn" not in sys.modules and importlib.util.find_spec("e3nn") is None:  # (other tests may install a stand-in)
        with pytest.raises(ImportError):
            sh_rotation_matrices_e3nn(torch.eye(3)[None], 2)
