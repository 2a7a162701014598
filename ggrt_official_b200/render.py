"""Host-side mirror of GGRt's render glue, for callers that do not have the reference tree.

Same names, arguments and maths as
/root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py (get_projection_matrix
:18-46, render_cuda :49-128, render_depth_cuda :227-269) and get_fov
(ggrt/geometry/projection.py:233-247), written against this package's rasterizer.  The
reference file itself also runs unmodified on top of the `diff_gaussian_rasterization`
shim; tests/test_reference_caller.py checks both produce identical rasterizer calls.
"""
from __future__ import annotations

from math import isqrt
from typing import Literal, Optional, Sequence

import torch
from torch import Tensor

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer

DepthRenderingMode = Literal["depth", "disparity", "relative_disparity", "log"]


def get_fov(intrinsics: Tensor) -> Tensor:
    """[b,3,3] normalised intrinsics -> [b,2] (fov_x, fov_y): angle between the un-projected
    mid-points of opposite image edges."""
    inv = intrinsics.inverse()

    def ray(p):
        vec = torch.tensor(p, dtype=torch.float32, device=intrinsics.device)
        vec = torch.einsum("bij,j->bi", inv, vec)
        return vec / vec.norm(dim=-1, keepdim=True)

    fov_x = (ray([0, 0.5, 1]) * ray([1, 0.5, 1])).sum(dim=-1).acos()
    fov_y = (ray([0.5, 0, 1]) * ray([0.5, 1, 1])).sum(dim=-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)


def get_projection_matrix(near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor, intrinsics: Tensor) -> Tensor:
    """Off-centre perspective matrix from the normalised intrinsics of batch element 0
    (the reference indexes intrinsics[0], cuda_splatting.py:39-42); x,y -> (-1,1), z -> (0,1)."""
    (b,) = near.shape
    out = torch.zeros((b, 4, 4), dtype=torch.float32, device=near.device)
    out[:, 0, 0] = 2 * near * intrinsics[0, 0, 0]
    out[:, 1, 1] = 2 * near * intrinsics[0, 1, 1]
    out[:, 0, 2] = 2 * intrinsics[0, 0, 2] - 1
    out[:, 1, 2] = 2 * intrinsics[0, 1, 2] - 1
    out[:, 3, 2] = 1
    out[:, 2, 2] = far / (far - near)
    out[:, 2, 3] = -(far * near) / (far - near)
    return out


def render_cuda(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple,
                background_color: Tensor, gaussian_means: Tensor, gaussian_covariances: Tensor,
                gaussian_sh_coefficients: Tensor, gaussian_opacities: Tensor, scale_invariant: bool = True,
                use_sh: bool = True, return_aux: bool = False, gaussian_aux: Optional[Tensor] = None):
    """[b] views of [b,g] Gaussians -> [b,3,h,w].  With return_aux also the per-view radii and the third
    output: the alpha-blended `gaussian_aux` [b,g] channel when given (differentiable), else the view depth."""
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    if scale_invariant:  # everything rescaled so that near == 1 (cuda_splatting.py:66-73)
        scale = 1 / near
        extrinsics = extrinsics.clone()
        extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scale[:, None]
        gaussian_covariances = gaussian_covariances * (scale[:, None, None, None] ** 2)
        gaussian_means = gaussian_means * scale[:, None, None]
        near = near * scale
        far = far * scale

    n = gaussian_sh_coefficients.shape[-1]
    degree = isqrt(n) - 1
    shs = gaussian_sh_coefficients.permute(0, 1, 3, 2).contiguous()  # b g xyz n -> b g n xyz

    b = extrinsics.shape[0]
    h, w = image_shape
    fov_x, fov_y = get_fov(intrinsics).unbind(dim=-1)
    tan_fov_x = (0.5 * fov_x).tan().tolist()  # one host sync for all views
    tan_fov_y = (0.5 * fov_y).tan().tolist()
    projection = get_projection_matrix(near, far, fov_x, fov_y, intrinsics).transpose(1, 2)
    view = extrinsics.inverse().transpose(1, 2)
    full = view @ projection
    row, col = torch.triu_indices(3, 3)

    images, radii_all, depths = [], [], []
    for i in range(b):
        mean_gradients = torch.zeros_like(gaussian_means[i], requires_grad=True)
        settings = GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=tan_fov_x[i], tanfovy=tan_fov_y[i], bg=background_color[i],
            scale_modifier=1.0, viewmatrix=view[i], projmatrix=full[i], sh_degree=degree,
            campos=extrinsics[i, :3, 3], prefiltered=False)
        image, radii, depth = GaussianRasterizer(settings)(
            means3D=gaussian_means[i], means2D=mean_gradients, shs=shs[i] if use_sh else None,
            colors_precomp=None if use_sh else shs[i, :, 0, :], opacities=gaussian_opacities[i, ..., None],
            cov3D_precomp=gaussian_covariances[i, :, row, col],
            aux_precomp=None if gaussian_aux is None else gaussian_aux[i])
        images.append(image)
        radii_all.append(radii)
        depths.append(depth)
    out = torch.stack(images)
    if return_aux:
        return out, torch.stack(radii_all), torch.stack(depths)
    return out


def depth_to_relative_disparity(depth: Tensor, near: Tensor, far: Tensor, eps: float = 1e-10) -> Tensor:
    """ggrt/model/pixelsplat/encoder/epipolar/conversions.py: 1 - (1/d - 1/far) / (1/near - 1/far)."""
    disp_near = 1 / (near + eps)
    disp_far = 1 / (far + eps)
    disp = 1 / (depth + eps)
    return 1 - (disp - disp_far) / (disp_near - disp_far + eps)


SH_C0 = 0.28209479177387814


def depth_channel(extrinsics: Tensor, gaussian_means: Tensor, near: Tensor, far: Tensor,
                  mode: DepthRenderingMode = "depth") -> Tensor:
    """The per-Gaussian value GGRt's depth pass blends: camera-space z (or its disparity / log variants,
    cuda_splatting.py:240-252) pushed through the degree-0 SH colour path, max(0, C0*z + 0.5)."""
    hom = torch.cat([gaussian_means, torch.ones_like(gaussian_means[..., :1])], dim=-1)
    fake = torch.einsum("bij,bgj->bgi", extrinsics.inverse(), hom)[..., 2]
    if mode == "disparity":
        fake = 1 / fake
    elif mode == "relative_disparity":
        fake = depth_to_relative_disparity(fake, near[:, None], far[:, None])
    elif mode == "log":
        fake = fake.minimum(near[:, None]).maximum(far[:, None]).log()
    return (SH_C0 * fake + 0.5).clamp(min=0.0)


def render_color_and_depth_cuda(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple,
                                background_color: Tensor, gaussian_means: Tensor, gaussian_covariances: Tensor,
                                gaussian_sh_coefficients: Tensor, gaussian_opacities: Tensor,
                                scale_invariant: bool = True, mode: DepthRenderingMode = "depth"):
    """render_cuda + render_depth_cuda in ONE rasterization: the depth pass blends a
    per-Gaussian scalar with exactly the colour pass's geometry and weights, so it rides along as the
    rasterizer's aux channel instead of costing a second preprocess + binning + render.
    Returns (color [b,3,h,w], depth [b,h,w]) equal to the two reference calls."""
    aux = depth_channel(extrinsics, gaussian_means, near, far, mode)
    color, _, depth = render_cuda(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                                  gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities,
                                  scale_invariant=scale_invariant, return_aux=True, gaussian_aux=aux)
    return color, depth


def render_views_fast(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple,
                      background_color: Tensor, gaussian_means: Tensor, gaussian_covariances: Tensor,
                      gaussian_sh_coefficients: Tensor, gaussian_opacities: Tensor, view_to_scene: Sequence[int],
                      scale_invariant: bool = True, depth_mode: Optional[DepthRenderingMode] = None):
    """Same result as render_cuda (+ render_depth_cuda when depth_mode is given), with the per-call copies of the
    reference glue removed:
      * the Gaussians are NOT replicated per view (`view_to_scene[i]` picks the scene of view i) and NOT rescaled,
        gathered or permuted in PyTorch: the rasterizer reads pixelSplat's own layout (means [g,3], covariances
        [g,3,3], harmonics [g,3,n]) and applies the scale-invariant 1/near factor itself (`layout` argument);
      * one host sync per call (tan(fov/2) and 1/near for all views together) instead of two per view;
      * colour and depth share one rasterization (aux channel).
    extrinsics / intrinsics / near / far / background_color are per view [b', ...]; the Gaussian tensors are per
    scene [s, g, ...].  Returns (color [b',3,h,w], depth [b',h,w] or None)."""
    nb = extrinsics.shape[0]
    h, w = image_shape
    n = gaussian_sh_coefficients.shape[-1]
    degree = isqrt(n) - 1
    aux = None
    if depth_mode is not None:  # camera-space z of the UNSCALED scene, per view (cuda_splatting.py:240-243)
        v2s = list(view_to_scene)
        per_view_means = gaussian_means if v2s == list(range(gaussian_means.shape[0])) else \
            gaussian_means[torch.as_tensor(v2s, device=gaussian_means.device)]
        aux = depth_channel(extrinsics, per_view_means, near, far, depth_mode)
    scale = 1 / near if scale_invariant else torch.ones_like(near)
    extr = extrinsics.clone()
    extr[..., :3, 3] = extr[..., :3, 3] * scale[:, None]
    near_s, far_s = near * scale, far * scale
    fov_x, fov_y = get_fov(intrinsics).unbind(dim=-1)
    host = torch.stack([(0.5 * fov_x).tan(), (0.5 * fov_y).tan(), scale]).tolist()  # the single host sync
    projection = get_projection_matrix(near_s, far_s, fov_x, fov_y, intrinsics).transpose(1, 2)
    view = extr.inverse().transpose(1, 2)
    full = view @ projection
    harm = gaussian_sh_coefficients.contiguous()
    cov = gaussian_covariances.contiguous()
    colors, depths = [], []
    for i in range(nb):
        s_ = view_to_scene[i]
        settings = GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=host[0][i], tanfovy=host[1][i], bg=background_color[i],
            scale_modifier=1.0, viewmatrix=view[i], projmatrix=full[i], sh_degree=degree, campos=extr[i, :3, 3],
            prefiltered=False)
        image, _, depth = GaussianRasterizer(settings)(
            means3D=gaussian_means[s_], means2D=None, shs=harm[s_], opacities=gaussian_opacities[s_],
            cov3D_precomp=cov[s_], aux_precomp=None if aux is None else aux[i],
            layout=dict(scene_scale=host[2][i], cov_full3x3=True, sh_channel_major=True))
        colors.append(image)
        depths.append(depth)
    return torch.stack(colors), (torch.stack(depths) if depth_mode is not None else None)


def camera_setup(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                 scale_invariant: bool = True, out: Optional[Tensor] = None) -> Tensor:
    """[n, 48] float32 camera blocks for n views -- viewmatrix [0:16], projmatrix [16:32], campos [32:35],
    {tanfovx, tanfovy, scene_scale} [35:38] -- computed by ONE kernel (ggrt_camera_setup) with no host read-back:
    the device-side twin of the ~60 small PyTorch operations and two .item() synchronisations per view of
    cuda_splatting.py:64-89,104-105.  The cameras are constants (the reference detaches poses)."""
    import ctypes as C

    from . import _cabi

    if not extrinsics.is_cuda:
        raise RuntimeError("extrinsics must be a CUDA tensor: the rasterizer has no CPU path")
    dev = extrinsics.device
    f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    E, K, nr, fr = f(extrinsics), f(intrinsics), f(near).reshape(-1), f(far).reshape(-1)
    n = E.shape[0]
    if tuple(E.shape) != (n, 4, 4) or tuple(K.shape) != (n, 3, 3) or nr.numel() != n or fr.numel() != n:
        raise ValueError(f"camera_setup: need extrinsics [n,4,4], intrinsics [n,3,3], near / far [n]; got "
                         f"{tuple(E.shape)}, {tuple(K.shape)}, {tuple(nr.shape)}, {tuple(fr.shape)}")
    if out is None:
        out = torch.empty((n, _cabi.CAMERA_FLOATS), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (n, _cabi.CAMERA_FLOATS) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise ValueError(f"camera_setup: out must be a contiguous float32 [{n},{_cabi.CAMERA_FLOATS}] tensor on {dev}")
    with torch.cuda.device(dev):
        sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        _cabi.check(_cabi.lib().ggrt_camera_setup(n, p(E), p(K), p(nr), p(fr), int(bool(scale_invariant)), p(out), sp),
                    "camera_setup")
    return out


def render_views_device(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple,
                        background_color: Tensor, gaussian_means: Tensor, gaussian_covariances: Tensor,
                        gaussian_sh_coefficients: Tensor, gaussian_opacities: Tensor, view_to_scene: Sequence[int],
                        scale_invariant: bool = True, depth: bool = False, streams: int = 1):
    """render_views_fast without any host synchronisation and with (almost) no PyTorch glue: the cameras of all views
    are set up by one kernel (camera_setup), the rasterizer reads tan(fov/2) and the scene scale from the device
    (GaussianRasterizationSettings.device_params), reads pixelSplat's own tensor layout, and -- with depth=True --
    evaluates GGRt's depth channel (mode "depth") inside its kernels (aux_mode=1), its gradient flowing straight
    into the means.  The views' kernels are enqueued back to back; nothing is read back.  Same outputs as
    render_cuda (+ render_depth_cuda) within the parity tolerance (the camera matrices may differ from the PyTorch
    glue's in the last ulp).
    `streams` > 1 (multi-view calls): the views are issued round-robin on that many CUDA streams (the caller's + cached
    side streams), forward AND -- because autograd runs a node's backward on its forward stream -- backward.  The views
    are independent (cuda_splatting.py:93-127), and the rasterizer alternates between latency-bound kernels (binning),
    issue-bound ones (the two render kernels) and HBM-bound ones (colour, per-Gaussian backward), so two views in
    flight fill each other's gaps instead of queueing 8 kernels per view one after another.
    Returns (color [n,3,h,w], depth [n,h,w] or None)."""
    nb = extrinsics.shape[0]
    h, w = image_shape
    degree = isqrt(gaussian_sh_coefficients.shape[-1]) - 1
    cams = camera_setup(extrinsics, intrinsics, near, far, scale_invariant)
    harm = gaussian_sh_coefficients.contiguous()
    cov = gaussian_covariances.contiguous()
    layout = dict(scene_scale=1.0, cov_full3x3=True, sh_channel_major=True)  # the scale itself is read on the device
    dev = extrinsics.device
    main = torch.cuda.current_stream(dev)
    pool = [main] + _side_streams(dev, max(0, min(int(streams), nb) - 1))
    for st in pool[1:]:
        st.wait_stream(main)  # the inputs (and the camera kernel) were produced on the caller's stream
    colors, depths = [], []
    for i in range(nb):
        s_ = view_to_scene[i]
        c = cams[i]
        st = pool[i % len(pool)]
        with torch.cuda.stream(st):
            settings = GaussianRasterizationSettings(
                image_height=h, image_width=w, tanfovx=0.0, tanfovy=0.0, bg=background_color[i], scale_modifier=1.0,
                viewmatrix=c[0:16], projmatrix=c[16:32], sh_degree=degree, campos=c[32:35], prefiltered=False,
                device_params=c[35:38], aux_mode=1 if depth else 0)
            image, _, d = GaussianRasterizer(settings)(
                means3D=gaussian_means[s_], means2D=None, shs=harm[s_], opacities=gaussian_opacities[s_],
                cov3D_precomp=cov[s_], layout=layout)
        if st is not main:  # the outputs are consumed on the caller's stream: keep the allocator from recycling them early
            image.record_stream(main)
            d.record_stream(main)
        colors.append(image)
        depths.append(d)
    for st in pool[1:]:
        main.wait_stream(st)
    return torch.stack(colors), (torch.stack(depths) if depth else None)


def _cabi_camera_floats() -> int:
    from . import _cabi

    return _cabi.CAMERA_FLOATS


_SIDE_STREAMS: dict = {}


def _side_streams(device, n: int) -> list:
    """n cached side streams of `device` for render_views_device(streams=...)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    have = _SIDE_STREAMS.setdefault(key, [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device=device))
    return have[:n]


def render_depth_cuda(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple,
                      gaussian_means: Tensor, gaussian_covariances: Tensor, gaussian_opacities: Tensor,
                      scale_invariant: bool = True, mode: DepthRenderingMode = "depth") -> Tensor:
    """Depth as a 3-channel degree-0 'colour' through render_cuda, then the channel mean
    (cuda_splatting.py:227-269); camera-space z is taken BEFORE the scale-invariant rescale."""
    hom = torch.cat([gaussian_means, torch.ones_like(gaussian_means[..., :1])], dim=-1)
    cam = torch.einsum("bij,bgj->bgi", extrinsics.inverse(), hom)
    fake = cam[..., 2]
    if mode == "disparity":
        fake = 1 / fake
    elif mode == "relative_disparity":
        fake = depth_to_relative_disparity(fake, near[:, None], far[:, None])
    elif mode == "log":
        fake = fake.minimum(near[:, None]).maximum(far[:, None]).log()
    b = fake.shape[0]
    result = render_cuda(extrinsics, intrinsics, near, far, image_shape,
                         torch.zeros((b, 3), dtype=fake.dtype, device=fake.device), gaussian_means,
                         gaussian_covariances, fake[..., None, None].expand(-1, -1, 3, 1), gaussian_opacities,
                         scale_invariant=scale_invariant)
    return result.mean(dim=1)
