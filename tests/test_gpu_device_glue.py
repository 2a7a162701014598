"""GPU tests of the sync-free render glue: the camera set-up kernel against the PyTorch mirror
of the reference glue (cuda_splatting.py:64-89, projection.py:233-247), and the in-kernel depth channel
(aux_mode = 1) against the explicit aux channel."""
import numpy as np
import pytest
import torch

from ggrt_official_b200 import GaussianRasterizationSettings, GaussianRasterizer
from ggrt_official_b200.decoder import DecoderSplattingCUDA, Gaussians
from ggrt_official_b200.render import camera_setup, depth_channel, get_fov, get_projection_matrix
from ggrt_official_b200.synthetic import make_scene, small_se3

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _views(n, seed=3):
    rng = np.random.default_rng(seed)
    sc = make_scene(2000, 96, 128, sh_degree=2, seed=seed, near=0.7, far=60.0)
    E = np.stack([sc.extrinsics.astype(np.float64) @ small_se3(rng, rot_deg=25.0, trans=0.6).astype(np.float64)
                  for _ in range(n)]).astype(np.float32)
    K = np.stack([sc.intrinsics for _ in range(n)]).astype(np.float32)
    K[:, 0, 2] += rng.uniform(-0.03, 0.03, n).astype(np.float32)  # off-centre principal points
    K[:, 0, 0] *= rng.uniform(0.8, 1.2, n).astype(np.float32)
    near = rng.uniform(0.5, 2.0, n).astype(np.float32)
    far = (near * rng.uniform(20, 200, n)).astype(np.float32)
    t = lambda a: torch.tensor(a, device=DEV)
    return sc, t(E), t(K), t(near), t(far)


@pytest.mark.parametrize("scale_invariant", [True, False])
def test_camera_setup_matches_the_pytorch_glue(scale_invariant):
    _, E, K, near, far = _views(5)
    cams = camera_setup(E, K, near, far, scale_invariant)
    assert cams.shape == (5, 48)
    scale = 1 / near if scale_invariant else torch.ones_like(near)
    extr = E.clone()
    extr[:, :3, 3] = extr[:, :3, 3] * scale[:, None]
    nr, fr = near * scale, far * scale
    fov_x, fov_y = get_fov(K).unbind(dim=-1)
    proj = get_projection_matrix(nr, fr, fov_x, fov_y, K).transpose(1, 2)
    view = extr.inverse().transpose(1, 2)
    full = view @ proj

    def close(a, b, tol=2e-6):
        return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))

    assert close(cams[:, 0:16].reshape(5, 4, 4), view)
    assert close(cams[:, 16:32].reshape(5, 4, 4), full, 5e-6)
    assert close(cams[:, 32:35], extr[:, :3, 3])
    assert close(cams[:, 35], (0.5 * fov_x).tan()) and close(cams[:, 36], (0.5 * fov_y).tan())
    assert close(cams[:, 37], scale) and close(cams[:, 38], nr) and close(cams[:, 39], fr)


def test_in_kernel_depth_channel_equals_explicit_aux():
    """aux_mode = 1 (depth channel evaluated inside the kernels, gradient into the means) against the same channel
    computed with PyTorch and passed as aux_precomp (gradient through autograd)."""
    sc, E, K, near, far = _views(1, seed=5)
    H, W = sc.image_shape
    t = lambda a: torch.tensor(np.asarray(a), device=DEV)
    cams = camera_setup(E, K, near, far, True)
    c = cams[0]
    means_a = t(sc.means).requires_grad_()
    means_b = t(sc.means).requires_grad_()
    cov, harm, opac = t(sc.covariances), t(sc.harmonics), t(sc.opacities)
    layout = dict(scene_scale=1.0, cov_full3x3=True, sh_channel_major=True)
    common = dict(image_height=H, image_width=W, tanfovx=0.0, tanfovy=0.0, bg=torch.zeros(3, device=DEV),
                  scale_modifier=1.0, viewmatrix=c[0:16], projmatrix=c[16:32], sh_degree=2, campos=c[32:35],
                  prefiltered=False, device_params=c[35:38])
    img_a, radii_a, dep_a = GaussianRasterizer(GaussianRasterizationSettings(aux_mode=1, **common))(
        means3D=means_a, means2D=None, shs=harm, opacities=opac, cov3D_precomp=cov, layout=layout)
    aux = depth_channel(E, means_b[None], near, far, "depth")[0]
    img_b, radii_b, dep_b = GaussianRasterizer(GaussianRasterizationSettings(**common))(
        means3D=means_b, means2D=None, shs=harm, opacities=opac, cov3D_precomp=cov, aux_precomp=aux, layout=layout)
    assert torch.equal(radii_a, radii_b) and torch.equal(img_a, img_b)
    assert float((dep_a - dep_b).abs().max()) <= 2e-6 * max(1.0, float(dep_b.abs().max()))
    wd = torch.randn(H, W, device=DEV)
    wc = torch.randn(3, H, W, device=DEV)
    ((img_a * wc).sum() + (dep_a * wd).sum()).backward()
    ((img_b * wc).sum() + (dep_b * wd).sum()).backward()
    assert float((means_a.grad - means_b.grad).abs().max()) <= 2e-5 * float(means_b.grad.abs().max())


def test_device_glue_decoder_equals_fast_glue_for_two_views():
    sc, E, K, near, far = _views(2, seed=7)
    H, W = sc.image_shape
    t = lambda a: torch.tensor(np.asarray(a), device=DEV)
    out = {}
    for name, kw in (("fast", dict(fast_glue=True)), ("device", dict(device_glue=True))):
        leaves = dict(means=t(sc.means)[None].requires_grad_(), covariances=t(sc.covariances)[None].requires_grad_(),
                      harmonics=t(sc.harmonics)[None].requires_grad_(), opacities=t(sc.opacities)[None].requires_grad_())
        r = DecoderSplattingCUDA(**kw)(Gaussians(**leaves), E[None], K[None], near[None], far[None], (H, W),
                                       depth_mode="depth")
        torch.manual_seed(0)
        wc, wd = torch.randn_like(r.color), torch.randn_like(r.depth)
        ((r.color * wc).sum() + (r.depth * wd).sum()).backward()
        out[name] = (r, leaves)
    (ra, la), (rb, lb) = out["fast"], out["device"]
    # the two glues differ in the last ulp of the camera matrices: a handful of threshold pixels may flip
    cerr = (ra.color - rb.color).abs().amax(dim=2)
    assert float((cerr > 1e-4).float().mean()) < 2e-3 and float(cerr.max()) < 2e-2
    derr = (ra.depth - rb.depth).abs() / max(1.0, float(ra.depth.abs().max()))
    assert float((derr > 1e-4).float().mean()) < 2e-3
    for k in la:
        ref, got = la[k].grad, lb[k].grad
        assert float((got - ref).abs().max()) <= 1e-3 * float(ref.abs().max()), k


@pytest.mark.parametrize("streams", [2, 3])
def test_views_on_several_streams_equal_the_sequential_views(streams):
    """render_views_device(streams=n): the views of a call issued round-robin on n CUDA streams (forward and, through
    autograd, backward) must give what the sequential loop gives -- images bit for bit (the forward is deterministic),
    gradients up to the summation order of the float atomics."""
    sc, E, K, near, far = _views(5, seed=11)
    H, W = sc.image_shape
    t = lambda a: torch.tensor(np.asarray(a), device=DEV)
    out = {}
    for name, kw in (("seq", dict(device_glue=True)), ("par", dict(device_glue=True, view_streams=streams))):
        leaves = dict(means=t(sc.means)[None].requires_grad_(), covariances=t(sc.covariances)[None].requires_grad_(),
                      harmonics=t(sc.harmonics)[None].requires_grad_(), opacities=t(sc.opacities)[None].requires_grad_())
        for _ in range(2):  # twice: the second pass reuses cached streams and allocator blocks
            for v in leaves.values():
                v.grad = None
            r = DecoderSplattingCUDA(**kw)(Gaussians(**leaves), E[None], K[None], near[None], far[None], (H, W),
                                           depth_mode="depth")
            torch.manual_seed(0)
            wc, wd = torch.randn_like(r.color), torch.randn_like(r.depth)
            ((r.color * wc).sum() + (r.depth * wd).sum()).backward()
        torch.cuda.synchronize()
        out[name] = (r, leaves)
    (ra, la), (rb, lb) = out["seq"], out["par"]
    assert torch.equal(ra.color, rb.color) and torch.equal(ra.depth, rb.depth)
    for k in la:
        ref, got = la[k].grad, lb[k].grad
        assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), k


@pytest.mark.parametrize("with_depth,streams", [(True, 2), (False, 1), (True, 3)])
def test_captured_views_equal_the_autograd_decoder_call(with_depth, streams):
    """graph.CapturedViews -- all views of a call as one CUDA-graph launch, compact per-view colour gradients and ONE
    SH-gradient merge -- against the same call through DecoderSplattingCUDA(device_glue=True) and autograd: images
    bit for bit, gradients (summed over the views) up to float summation order; also after several replays and after
    the cameras were changed in place."""
    from ggrt_official_b200.graph import CapturedViews

    sc, E, K, near, far = _views(4, seed=21)
    H, W = sc.image_shape
    t = lambda a: torch.tensor(np.asarray(a), device=DEV)
    torch.manual_seed(1)
    wc = torch.randn(4, 3, H, W, device=DEV)
    wd = torch.randn(4, H, W, device=DEV) if with_depth else None
    bg = torch.zeros(4, 3, device=DEV)
    cap = CapturedViews(E, K, near, far, (H, W), bg, t(sc.means), t(sc.covariances), t(sc.harmonics), t(sc.opacities),
                        grad_color=wc, grad_depth=wd, streams=streams)

    def reference(E_):
        leaves = dict(means=t(sc.means)[None].requires_grad_(), covariances=t(sc.covariances)[None].requires_grad_(),
                      harmonics=t(sc.harmonics)[None].requires_grad_(), opacities=t(sc.opacities)[None].requires_grad_())
        r = DecoderSplattingCUDA(device_glue=True)(Gaussians(**leaves), E_[None], K[None], near[None], far[None], (H, W),
                                                   depth_mode="depth" if with_depth else None)
        loss = (r.color[0] * wc).sum()
        if with_depth:
            loss = loss + (r.depth[0] * wd).sum()
        loss.backward()
        return r, {k: v.grad[0] for k, v in leaves.items()}

    for trial in range(2):
        if trial == 1:  # move the cameras in place: the captured camera kernel must pick the new poses up
            rng = np.random.default_rng(5)
            E2 = torch.tensor(np.stack([E[i].cpu().numpy().astype(np.float64) @ small_se3(rng, rot_deg=5.0, trans=0.1)
                                        for i in range(4)]).astype(np.float32), device=DEV)
            cap.cam_in[0].copy_(E2)
        for _ in range(3):
            cap.replay()
        cap.check()
        r, g = reference(cap.cam_in[0].clone())
        assert torch.equal(cap.color, r.color[0])
        if with_depth:
            assert torch.equal(cap.depth, r.depth[0])
        got = cap.grads
        for k, name in (("means", "dmeans"), ("covariances", "dcovariances"), ("harmonics", "dharmonics"),
                        ("opacities", "dopacities")):
            ref = g[k]
            assert got[name].shape == ref.shape, (name, got[name].shape, ref.shape)
            assert float((got[name] - ref).abs().max()) <= 3e-5 * float(ref.abs().max()), name
