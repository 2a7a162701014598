"""Fused Gaussian construction: pixelSplat's `GaussianAdapter` on one forward and one backward CUDA kernel.

Mirror of /root/reference/ggrt/model/pixelsplat/encoder/common/gaussian_adapter.py: same
`GaussianAdapterCfg`, `Gaussians` dataclass, `forward(extrinsics, intrinsics, coordinates, depths, opacities,
raw_gaussians, image_shape, eps)` signature and results, called as at encoder_epipolar.py:221-229 with the batch
structure "b v r srf spp" (extrinsics / intrinsics per view, coordinates and raw features per ray, depths and
opacities per Gaussian).  The reference evaluates ~30 PyTorch kernels (sigmoid, quaternion -> matrix, four 3x3
matmuls, ray unprojection, SH mask + per-degree Wigner-D einsums) with [G,3,3] / [G,3,K] intermediates; here
`ggrt_adapter_forward` / `ggrt_adapter_backward` (csrc/adapter.cu) read the raw features once and write the
rasterizer's inputs once.  No CPU path.

SH rotation: the reference calls `rotate_sh` (ggrt/misc/sh_rotation.py:10-29), i.e. e3nn's `wigner_D` of the
view's camera-to-world rotation, one (2l+1)x(2l+1) block per degree.  Those per-VIEW matrices are tiny and are
computed on the host side -- by e3nn when it is installed (`sh_rotation_matrices_e3nn`, the same two calls the
reference makes), or supplied by the caller (`sh_rotations=` [.., K, K]) -- and applied per Gaussian in the kernel.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from math import prod
from typing import Callable, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _cabi


@dataclass
class Gaussians:  # gaussian_adapter.py:13-20
    means: Tensor
    covariances: Tensor
    scales: Tensor
    rotations: Tensor
    harmonics: Tensor
    opacities: Tensor


@dataclass
class GaussianAdapterCfg:  # gaussian_adapter.py:23-27
    gaussian_scale_min: float
    gaussian_scale_max: float
    sh_degree: int


def sh_rotation_matrices_e3nn(rotations: Tensor, sh_degree: int) -> Tensor:
    """[.., 3, 3] rotations -> [.., K, K] block-diagonal SH rotation, exactly the matrices `rotate_sh` applies
    (sh_rotation.py:18-22).  Needs e3nn (a dependency of the reference)."""
    from e3nn.o3 import matrix_to_angles, wigner_D  # noqa: PLC0415 - optional dependency of the reference

    alpha, beta, gamma = matrix_to_angles(rotations)
    K = (sh_degree + 1) ** 2
    out = rotations.new_zeros(*rotations.shape[:-2], K, K)
    for degree in range(sh_degree + 1):
        blk = wigner_D(torch.tensor(degree).to(rotations.device), alpha, beta, gamma).type(rotations.dtype)
        out[..., degree ** 2: (degree + 1) ** 2, degree ** 2: (degree + 1) ** 2] = blk
    return out


def _c(t: Tensor) -> Tensor:
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _params(V, rays, spp, deg, h, w, smin, smax, eps) -> "_cabi.AdapterParams":
    p = _cabi.AdapterParams()
    p.num_views, p.rays_per_view, p.samples_per_ray, p.sh_degree = V, rays, spp, deg
    p.image_height, p.image_width = h, w
    p.scale_min, p.scale_max, p.eps = float(smin), float(smax), float(eps)
    return p


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class _AdapterFunction(torch.autograd.Function):
    """(coordinates [R,2], depths [G], raw [R,7+3K]) -> (means [G,3], covariances [G,3,3], harmonics [G,3,K],
    scales [G,3], rotations [G,4]); cameras and SH rotations are constants."""

    @staticmethod
    def forward(ctx, coordinates, depths, raw, extrinsics, intrinsics, sh_rot, meta):
        V, rays, spp, deg, h, w, smin, smax, eps = meta
        L = _cabi.lib()
        dev = raw.device
        G, K = V * rays * spp, (deg + 1) ** 2
        f32 = dict(dtype=torch.float32, device=dev)
        means, cov = torch.empty((G, 3), **f32), torch.empty((G, 3, 3), **f32)
        harm, scales, rots = torch.empty((G, 3, K), **f32), torch.empty((G, 3), **f32), torch.empty((G, 4), **f32)
        p = _params(*meta)
        with torch.cuda.device(dev):
            sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(L.ggrt_adapter_forward(C.byref(p), _ptr(extrinsics), _ptr(intrinsics), _ptr(sh_rot),
                                               _ptr(coordinates), _ptr(depths), _ptr(raw), _ptr(means), _ptr(cov),
                                               _ptr(harm), _ptr(scales), _ptr(rots), sp), "adapter_forward")
        ctx.save_for_backward(coordinates, depths, raw, extrinsics, intrinsics, sh_rot)
        ctx.meta = meta
        ctx.mark_non_differentiable(scales, rots)  # only exported / visualised by the reference (detached there too)
        ctx.set_materialize_grads(False)
        return means, cov, harm, scales, rots

    @staticmethod
    def backward(ctx, g_means, g_cov, g_harm, _g_scales=None, _g_rots=None):
        coordinates, depths, raw, extrinsics, intrinsics, sh_rot = ctx.saved_tensors
        if g_means is None and g_cov is None and g_harm is None:
            return (None,) * 7
        L = _cabi.lib()
        dev = raw.device
        g_means = None if g_means is None else _c(g_means)
        g_cov = None if g_cov is None else _c(g_cov)
        g_harm = None if g_harm is None else _c(g_harm)
        d_coords, d_depths, d_raw = torch.empty_like(coordinates), torch.empty_like(depths), torch.empty_like(raw)
        p = _params(*ctx.meta)
        with torch.cuda.device(dev):
            sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(L.ggrt_adapter_backward(C.byref(p), _ptr(extrinsics), _ptr(intrinsics), _ptr(sh_rot),
                                                _ptr(coordinates), _ptr(depths), _ptr(raw), _ptr(g_means), _ptr(g_cov),
                                                _ptr(g_harm), _ptr(d_coords), _ptr(d_depths), _ptr(d_raw), sp),
                        "adapter_backward")
        return d_coords, d_depths, d_raw, None, None, None, None


class GaussianAdapter(nn.Module):
    cfg: GaussianAdapterCfg

    def __init__(self, cfg: GaussianAdapterCfg, sh_rotation_fn: Optional[Callable[[Tensor, int], Tensor]] = None):
        """`sh_rotation_fn(c2w_rotations [.., 3, 3], sh_degree) -> [.., K, K]`; default: e3nn, as the reference."""
        super().__init__()
        self.cfg = cfg
        if not 0 <= cfg.sh_degree <= 4:
            raise ValueError(f"sh_degree {cfg.sh_degree} outside 0..4")
        self.sh_rotation_fn = sh_rotation_fn or sh_rotation_matrices_e3nn
        # kept for state-dict / attribute compatibility; the kernel has the same constants built in
        self.register_buffer("sh_mask", torch.ones((self.d_sh,), dtype=torch.float32), persistent=False)
        for degree in range(1, cfg.sh_degree + 1):
            self.sh_mask[degree ** 2: (degree + 1) ** 2] = 0.1 * 0.25 ** degree

    def forward(self, extrinsics: Tensor, intrinsics: Tensor, coordinates: Tensor, depths: Tensor, opacities: Tensor,
                raw_gaussians: Tensor, image_shape: Tuple[int, int], eps: float = 1e-8,
                sh_rotations: Optional[Tensor] = None) -> Gaussians:
        if not raw_gaussians.is_cuda:
            raise RuntimeError("raw_gaussians must be a CUDA tensor: the fused adapter has no CPU path")
        if depths.dim() < 3:
            raise ValueError("depths must be [..., ray, surface, sample] as at encoder_epipolar.py:221-229")
        *lead, r, srf, spp = depths.shape
        lead = tuple(lead)
        V, rays, K = prod(lead) if lead else 1, r * srf, self.d_sh
        full = lead + (r, srf)

        def per_view(t, tail, name):
            want = lead + (1, 1, 1) + tail
            if t.dim() != len(want) or any(a not in (1, b) for a, b in zip(t.shape, want)) or tuple(t.shape[-len(tail):]) != tail:
                raise NotImplementedError(f"{name} must be per view, broadcastable to {want}; got {tuple(t.shape)}")
            return _c(t.expand(want)).reshape(V, *tail)

        def per_ray(t, tail, name):
            want = full + (1,) + tail
            if t.dim() != len(want) or any(a not in (1, b) for a, b in zip(t.shape, want)) or tuple(t.shape[-len(tail):]) != tail:
                raise NotImplementedError(f"{name} must be per ray (shared by the samples), broadcastable to {want}; "
                                          f"got {tuple(t.shape)}")
            return _c(t.expand(want)).reshape(V * rays, *tail)

        extr = per_view(extrinsics, (4, 4), "extrinsics")
        intr = per_view(intrinsics, (3, 3), "intrinsics")
        coords = per_ray(coordinates, (2,), "coordinates")
        raw = per_ray(raw_gaussians, (7 + 3 * K,), "raw_gaussians")
        dep = _c(depths).reshape(-1)
        with torch.no_grad():  # cameras are constants here, exactly as rotate_sh's angles are in the reference
            if sh_rotations is None and K > 1:
                sh_rotations = self.sh_rotation_fn(extr[:, :3, :3], self.cfg.sh_degree)
            rot = None if sh_rotations is None else _c(sh_rotations.detach().expand(*lead, K, K)
                                                       if sh_rotations.dim() == len(lead) + 2 else sh_rotations.detach()
                                                       ).reshape(V, K, K)
        h, w = image_shape
        meta = (V, rays, spp, self.cfg.sh_degree, int(h), int(w), self.cfg.gaussian_scale_min,
                self.cfg.gaussian_scale_max, eps)
        means, cov, harm, scales, rots = _AdapterFunction.apply(coords, dep, raw, extr.detach(), intr.detach(), rot, meta)
        shape = tuple(depths.shape)
        return Gaussians(
            means=means.view(*shape, 3),
            covariances=cov.view(*shape, 3, 3),
            harmonics=harm.view(*shape, 3, K),
            opacities=opacities,
            scales=scales.view(*shape, 3),
            rotations=rots.view(*shape, 4),
        )

    def get_scale_multiplier(self, intrinsics: Tensor, pixel_size: Tensor, multiplier: float = 0.1) -> Tensor:
        """gaussian_adapter.py:98-109 (kept for API completeness; the kernel computes it per view)."""
        xy = multiplier * torch.einsum("...ij,...j->...i", intrinsics[..., :2, :2].inverse(), pixel_size)
        return xy.sum(dim=-1)

    @property
    def d_sh(self) -> int:
        return (self.cfg.sh_degree + 1) ** 2

    @property
    def d_in(self) -> int:
        return 7 + 3 * self.d_sh
