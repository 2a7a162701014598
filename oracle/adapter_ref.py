"""CPU/torch restatement of pixelSplat's Gaussian adapter.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/ggrt/model/pixelsplat/encoder/common/gaussian_adapter.py:48-96 (forward, scale
multiplier :98-109), gaussians.py:8-44 (quaternion_to_matrix / build_covariance), ggrt/geometry/projection.py:74-114
(unproject / get_world_rays) and ggrt/misc/sh_rotation.py:10-29 (rotate_sh) with the per-degree Wigner-D blocks
given explicitly.  Pinned: tests/test_adapter_oracle.py checks it against golden vectors produced by the
UNMODIFIED reference classes (tools/make_golden_adapter.py).  Differentiable (autograd), any float dtype.
"""
from __future__ import annotations

import torch
from torch import Tensor


def quaternion_to_matrix(q: Tensor, eps: float = 1e-8) -> Tensor:  # gaussians.py:8-31, xyzw order
    i, j, k, r = torch.unbind(q, dim=-1)
    two_s = 2 / ((q * q).sum(dim=-1) + eps)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(*q.shape[:-1], 3, 3)


def adapter_forward(extrinsics: Tensor, intrinsics: Tensor, coordinates: Tensor, depths: Tensor, raw: Tensor,
                    image_shape, sh_degree: int, scale_min: float, scale_max: float, sh_rotations: Tensor = None,
                    eps: float = 1e-8) -> dict:
    """extrinsics [V,4,4], intrinsics [V,3,3], coordinates [V,R,2], depths [V,R,S], raw [V,R,7+3K],
    sh_rotations [V,K,K] or None -> means [V,R,S,3], covariances [V,R,S,3,3], harmonics [V,R,S,3,K], scales, rotations."""
    K = (sh_degree + 1) ** 2
    V, R, S = depths.shape
    s_raw, q_raw, sh = raw.split((3, 4, 3 * K), dim=-1)
    h, w = image_shape
    scales = scale_min + (scale_max - scale_min) * s_raw.sigmoid()                       # [V,R,3]
    pixel_size = 1 / torch.tensor((w, h), dtype=raw.dtype)
    mult = (0.1 * torch.einsum("vij,j->vi", intrinsics[:, :2, :2].inverse(), pixel_size)).sum(-1)  # [V]
    scales = scales[:, :, None, :] * depths[..., None] * mult[:, None, None, None]      # [V,R,S,3]
    q = q_raw / (q_raw.norm(dim=-1, keepdim=True) + eps)                                 # [V,R,4]
    rot = quaternion_to_matrix(q)[:, :, None]                                            # [V,R,1,3,3]
    m = rot * scales[..., None, :]                                                       # R diag(s)
    cov = m @ m.transpose(-1, -2)
    c2w = extrinsics[:, None, None, :3, :3]
    cov = c2w @ cov @ c2w.transpose(-1, -2)
    xy1 = torch.cat((coordinates, torch.ones_like(coordinates[..., :1])), -1)            # [V,R,3]
    d = torch.einsum("vij,vrj->vri", intrinsics.inverse(), xy1)
    d = d / d.norm(dim=-1, keepdim=True)
    d = torch.einsum("vij,vrj->vri", extrinsics[:, :3, :3], d)
    means = extrinsics[:, None, None, :3, 3] + d[:, :, None, :] * depths[..., None]
    mask = torch.ones(K, dtype=raw.dtype)
    for deg in range(1, sh_degree + 1):
        mask[deg ** 2: (deg + 1) ** 2] = 0.1 * 0.25 ** deg
    sh = sh.reshape(V, R, 3, K) * mask
    if sh_rotations is not None:
        out = []
        for deg in range(sh_degree + 1):
            a, b = deg ** 2, (deg + 1) ** 2
            out.append(torch.einsum("vij,vrcj->vrci", sh_rotations[:, a:b, a:b], sh[..., a:b]))
        sh = torch.cat(out, dim=-1)
    harm = sh[:, :, None].expand(V, R, S, 3, K)
    return dict(means=means, covariances=cov, harmonics=harm, scales=scales,
                rotations=q[:, :, None].expand(V, R, S, 4))
