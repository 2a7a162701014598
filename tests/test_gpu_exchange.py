"""GPU tests of the compact gradient exchange kernels (single device, several "virtual" views).

The multi-GPU path exchanges [P,3] colour gradients instead of [P,K,3] SH gradients and rebuilds
dL/dsh = sum_v basis(dir_v) (x) dL/drgb_v locally (ggrt_raster_backward compact mode +
ggrt_raster_sh_gradient_merge).  Here all views live on one GPU: the merged result must equal the
sum of the full per-view backward passes, and -- per view -- the oracle's SH gradient.
"""
import copy

import numpy as np
import pytest
import torch

from ggrt_official_b200 import rasterizer as R
from ggrt_official_b200.synthetic import image_gradient, small_se3
from ggrt_official_b200.view_parallel import CompactGradientExchange
from oracle import c_oracle as co
from tests import gpu_util as G
from tests.helpers import oracle_camera, small_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, rel=2e-5):
    return float((a - b).abs().max()) <= rel * float(b.abs().max()) + 1e-30


def _views(P, H, W, deg, n_views, seed):
    """n_views RasterInputs of the same Gaussians seen from perturbed cameras."""
    from ggrt_official_b200.synthetic import to_raster_inputs

    sc, ri0 = small_case(P, H, W, deg, bg=(0.2, 0.1, 0.0), seed=seed, cov_scale=4.0)
    rng = np.random.default_rng(seed + 100)
    out = [ri0]
    for _ in range(n_views - 1):
        s2 = copy.copy(sc)
        s2.extrinsics = (sc.extrinsics.astype(np.float64) @ small_se3(rng).astype(np.float64)).astype(np.float32)
        ri = to_raster_inputs(s2, bg=(0.2, 0.1, 0.0))
        ri.cov3D = ri0.cov3D
        out.append(ri)
    return out


def _forward(ri, dev=DEV, layout=None, shs=None):
    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    sh = t(ri.shs) if shs is None else shs
    return R.forward_raw(t(ri.means3D), sh, None, t(ri.opacities), t(ri.cov3D), G.settings_from(ri, dev),
                         layout=layout)


@pytest.mark.parametrize("P,H,W,deg,n_views", [(3000, 64, 80, 4, 3), (1001, 48, 64, 2, 2), (640, 32, 32, 0, 5),
                                               (2048, 40, 56, 3, 9), (777, 32, 48, 1, 16)])
def test_merge_of_compact_gradients_equals_sum_of_full_backwards(P, H, W, deg, n_views):
    views = _views(P, H, W, deg, n_views, seed=11 + deg)
    drgb, cams, full_sum, small_full, small_compact = [], [], None, None, None
    for k, ri in enumerate(views):
        g_img = torch.tensor(image_gradient(H, W, seed=50 + k) * (3 * H * W), device=DEV)
        st = _forward(ri)
        full = R.backward_raw(st, g_img)
        comp = R.backward_raw(st, g_img, compact=True)
        assert comp["dsh"] is None and comp["dcolors"].shape == (P, 3)
        # everything but the SH gradient is the same computation (up to the order of the float atomics)
        for name in ("dmeans3D", "dcov3D", "dopacity", "dmeans2D"):
            assert _close(comp[name], full[name]), name
        # per view: dsh == basis(dir) (x) dcolors, against the oracle's SH gradient
        cam = oracle_camera(ri)
        f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
        b = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g_img.cpu().numpy(), sh=ri.shs)
        one = R.sh_gradient_merge(st["call"].means3D, deg, [comp["dcolors"]], [st["call"].campos])
        ref = b["dsh"]
        assert np.abs(one.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()
        assert np.abs(one.cpu().numpy() - full["dsh"].cpu().numpy()).max() <= 1e-5 * np.abs(ref).max()
        drgb.append(comp["dcolors"])
        cams.append(st["call"].campos)
        full_sum = full["dsh"].double() if full_sum is None else full_sum + full["dsh"].double()
    merged = R.sh_gradient_merge(torch.tensor(views[0].means3D, device=DEV), deg, drgb, cams)
    torch.cuda.synchronize()
    scale = float(full_sum.abs().max())
    assert scale > 0
    assert float((merged.double() - full_sum).abs().max()) <= 2e-5 * scale


def test_merge_channel_major_scaled_layout_and_addresses():
    """Channel-major SH, a scene scale, raw device addresses instead of tensors, unaligned output."""
    P, H, W, deg = 1500, 48, 64, 4
    (ri,) = _views(P, H, W, deg, 1, seed=5)
    t = lambda a: torch.tensor(np.asarray(a), device=DEV)
    scale = 0.5
    lay = dict(scene_scale=scale, cov_full3x3=False, sh_channel_major=True)
    sh_cm = t(ri.shs).permute(0, 2, 1).contiguous()
    rs = G.settings_from(ri, DEV)
    st = R.forward_raw(t(ri.means3D) / scale, sh_cm, None, t(ri.opacities), t(ri.cov3D) / (scale * scale), rs,
                       layout=lay)
    g_img = t(image_gradient(H, W, seed=3) * (3 * H * W))
    full = R.backward_raw(st, g_img)
    comp = R.backward_raw(st, g_img, compact=True)
    backing = torch.zeros(P * 75 + 1, device=DEV)
    out = backing[1:].view(P, 3, 25)  # 4-byte aligned only: the kernel's non-TMA write path
    got = R.sh_gradient_merge(st["call"].means3D, deg, [comp["dcolors"].data_ptr()], [st["call"].campos.data_ptr()],
                              out=out, layout=lay)
    torch.cuda.synchronize()
    ref = full["dsh"]
    assert got.shape == ref.shape == (P, 3, 25)
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_single_rank_exchange_object_and_argument_checks():
    P, H, W, deg = 1200, 32, 48, 4
    (ri,) = _views(P, H, W, deg, 1, seed=8)
    st = _forward(ri)
    g_img = torch.tensor(image_gradient(H, W, seed=1) * (3 * H * W), device=DEV)
    full = R.backward_raw(st, g_img)
    ex = CompactGradientExchange(P, deg, DEV)
    got = ex.run(st, g_img)
    torch.cuda.synchronize()
    for k in ("dmeans3D", "dcov3D", "dopacity", "dmeans2D"):
        assert _close(got[k], full[k]), k
    assert float((got["dsh"] - full["dsh"]).abs().max()) <= 1e-5 * float(full["dsh"].abs().max())
    with pytest.raises(ValueError):
        R.sh_gradient_merge(st["call"].means3D, deg, [], [])
    with pytest.raises(ValueError):
        R.sh_gradient_merge(st["call"].means3D, deg, [got["dsh"]], [st["call"].campos])  # wrong size
    with pytest.raises(ValueError):
        st2 = R.forward_raw(st["call"].means3D, None, torch.rand(P, 3, device=DEV), st["call"].opacities,
                            st["call"].cov3D, G.settings_from(ri, DEV))
        R.backward_raw(st2, g_img, compact=True)


@pytest.mark.parametrize("P", [1200, 1001])  # 1001: ragged last slab, scalar tail stores
def test_color_sinks_receive_gradients_and_campos(P):
    """Push model: the backward kernel itself replicates the compact colour gradients (+ the campos row) into
    several [P+1,3] buffers (here two local ones; on the multi-GPU path they are the peers' symmetric buffers)."""
    H, W, deg = 32, 48, 4
    (ri,) = _views(P, H, W, deg, 1, seed=9)
    st = _forward(ri)
    g_img = torch.tensor(image_gradient(H, W, seed=2) * (3 * H * W), device=DEV)
    comp = R.backward_raw(st, g_img, compact=True)
    n = (3 * (P + 1) + 3) // 4 * 4
    sinks = [torch.full((n,), float("nan"), device=DEV) for _ in range(2)]
    out = R.backward_raw(st, g_img, color_sinks={"ptrs": [b.data_ptr() for b in sinks], "multimem": False})
    torch.cuda.synchronize()
    assert out["dcolors"] is None and out["dsh"] is None
    for b in sinks:
        assert _close(b[: 3 * P].view(P, 3), comp["dcolors"])
        assert torch.equal(b[3 * P: 3 * P + 3], st["call"].campos)
    assert _close(out["dmeans3D"], comp["dmeans3D"])
    with pytest.raises(RuntimeError, match="aligned"):
        R.backward_raw(st, g_img, color_sinks={"ptrs": [sinks[0].data_ptr() + 4], "multimem": False})
