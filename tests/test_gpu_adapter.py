"""GPU parity of the fused Gaussian adapter (csrc/adapter.cu through the C ABI) against the golden vectors of the
unmodified reference GaussianAdapter and, on larger random shapes, against the pinned oracle.
Tolerances: outputs 1e-5 relative to the tensor's max (float32 vs the reference's float32), gradients 1e-4."""
import pytest
import torch

from ggrt_official_b200.adapter import GaussianAdapter, GaussianAdapterCfg
from tests.adapter_util import CASES, load_case, oracle_on_case, rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(c, dev=DEV):
    cfg = GaussianAdapterCfg(gaussian_scale_min=c["smin"], gaussian_scale_max=c["smax"], sh_degree=c["deg"])
    ad = GaussianAdapter(cfg).to(dev)
    b, v = c["depths"].shape[:2]
    K = (c["deg"] + 1) ** 2
    coords = c["coordinates"].to(dev).requires_grad_()
    depths = c["depths"].to(dev).requires_grad_()
    raw = c["raw"].to(dev).requires_grad_()
    out = ad(c["extrinsics"].to(dev)[:, :, None, None, None], c["intrinsics"].to(dev)[:, :, None, None, None], coords, depths,
             c["opacities"].to(dev), raw, c["image_shape"], sh_rotations=c["blocks"].to(dev).reshape(b, v, K, K))
    loss = sum((getattr(out, n) * c["up"][n].to(dev)).sum() for n in ("means", "covariances", "harmonics"))
    loss.backward()
    torch.cuda.synchronize()
    return out, dict(coordinates=coords.grad, depths=depths.grad, raw=raw.grad)


@pytest.mark.parametrize("name", CASES)
def test_matches_the_reference_golden(name):
    c = load_case(name)
    out, grads = _run(c)
    for k in ("means", "covariances", "harmonics", "scales", "rotations"):
        got = getattr(out, k)
        assert tuple(got.shape) == tuple(c["out"][k].shape), k
        assert rel(got, c["out"][k]) < 1e-5, (k, rel(got, c["out"][k]))
    assert torch.equal(out.opacities.cpu(), c["out"]["opacities"])
    for k in ("coordinates", "depths", "raw"):
        assert rel(grads[k], c["grad"][k]) < 1e-4, (k, rel(grads[k], c["grad"][k]))


@pytest.mark.parametrize("b,v,h,w,srf,spp,deg", [(1, 4, 24, 32, 1, 3, 4), (1, 2, 17, 13, 1, 1, 0), (2, 2, 9, 11, 2, 5, 3),
                                                  (1, 1, 8, 8, 1, 40, 1)])
def test_matches_the_oracle_on_random_shapes(b, v, h, w, srf, spp, deg):
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location(
        "make_golden_adapter", Path(__file__).resolve().parent.parent / "tools" / "make_golden_adapter.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    case = mg.make_case(seed=100 + deg, b=b, v=v, h=h, w=w, srf=srf, spp=spp, deg=deg)
    c = dict(case, out=None, grad=None)
    ref_out, ref_grads = oracle_on_case(c)
    out, grads = _run(c)
    for k in ("means", "covariances", "harmonics", "scales", "rotations"):
        assert rel(getattr(out, k), ref_out[k]) < 1e-5, (k, rel(getattr(out, k), ref_out[k]))
    for k in ("coordinates", "depths", "raw"):
        assert rel(grads[k], ref_grads[k]) < 1e-4, (k, rel(grads[k], ref_grads[k]))


def test_feeds_the_rasterizer_and_rejects_bad_inputs():
    """Adapter output -> rasterizer (pixelSplat layout: [G,3,3] covariances, channel-major harmonics) runs end to end
    with gradients back to the raw features; CPU tensors and per-sample raw features are refused."""
    from ggrt_official_b200 import GaussianRasterizationSettings, GaussianRasterizer

    H, W = 32, 48
    cfg = GaussianAdapterCfg(0.5, 15.0, 2)
    ad = GaussianAdapter(cfg, sh_rotation_fn=lambda rot, deg: torch.eye(9, device=rot.device).expand(rot.shape[0], 9, 9))
    g = torch.Generator().manual_seed(0)
    r = H * W
    extr = torch.eye(4, device=DEV)[None, None]  # camera at the origin looking down +z (the synthetic camera's frame)
    intr = torch.tensor([[0.9, 0, 0.5], [0, 1.2, 0.5], [0, 0, 1.0]], device=DEV)[None, None]
    yy, xx = torch.meshgrid((torch.arange(H) + 0.5) / H, (torch.arange(W) + 0.5) / W, indexing="ij")
    coords = torch.stack((xx, yy), -1).reshape(1, 1, r, 1, 1, 2).to(DEV)
    depths = (2.0 + torch.rand(1, 1, r, 1, 2, generator=g)).to(DEV).requires_grad_()
    raw = torch.randn(1, 1, r, 1, 1, 7 + 27, generator=g).to(DEV).requires_grad_()
    opac = torch.full((1, 1, r, 1, 2), 0.3, device=DEV)
    gs = ad(extr[:, :, None, None, None], intr[:, :, None, None, None], coords, depths, opac, raw, (H, W))
    P = r * 2
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=0.5 / 0.9, tanfovy=0.5 / 1.2, bg=torch.zeros(3, device=DEV), scale_modifier=1.0,
        viewmatrix=torch.eye(4, device=DEV),  # camera = world; projmatrix = (P of cuda_splatting.py:18-46)^T, near 0.1
        projmatrix=torch.tensor([[2 * 0.9, 0, 0, 0], [0, 2 * 1.2, 0, 0], [0, 0, 1.0, 1.0], [0, 0, -0.1, 0]], device=DEV),
        sh_degree=2, campos=torch.zeros(3, device=DEV), prefiltered=False)
    means2D = torch.zeros(P, 3, device=DEV, requires_grad=True)
    color, radii, _ = GaussianRasterizer(rs)(
        means3D=gs.means.reshape(P, 3), means2D=means2D, opacities=gs.opacities.reshape(P, 1),
        shs=gs.harmonics.reshape(P, 3, 9), cov3D_precomp=gs.covariances.reshape(P, 3, 3),
        layout=dict(scene_scale=1.0, cov_full3x3=True, sh_channel_major=True))
    color.square().sum().backward()
    torch.cuda.synchronize()
    assert int((radii > 0).sum()) > P // 2
    assert torch.isfinite(raw.grad).all() and float(raw.grad.abs().max()) > 0
    assert torch.isfinite(depths.grad).all() and float(depths.grad.abs().max()) > 0
    with pytest.raises(RuntimeError, match="no CPU path"):
        ad(extr.cpu()[:, :, None, None, None], intr.cpu()[:, :, None, None, None], coords.cpu(), depths.detach().cpu(),
           opac.cpu(), raw.detach().cpu(), (H, W))
    with pytest.raises(NotImplementedError, match="per ray"):
        ad(extr[:, :, None, None, None], intr[:, :, None, None, None], coords, depths, opac,
           raw.detach().expand(1, 1, r, 1, 2, 34), (H, W))
