"""Synthetic scenes with the statistics of GGRt's predicted Gaussians.

Numpy only (runs on the bench host and in CPU tests).  Distributions mirror
/root/reference/ggrt/model/pixelsplat/encoder/common/gaussian_adapter.py:60-96
(covariance = R diag(s^2) R^T, scale proportional to depth and pixel size,
`0.5 + 14.5*sigmoid` multiplier), depth_predictor_monocular.py:63-68 (uniform
disparity), encoder_epipolar.py:195 (opacity <= 1/gaussians_per_pixel = 1/3) and
gaussian_adapter.py:45-46 (SH band l scaled by 0.1*0.25^l).  Seed 3407 is the
reference's default (configs/pretrain_ggrt_stable.yaml:13).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEED = 3407


@dataclass
class Scene:
    """Decoder-level inputs, b = v = 1 squeezed: what DecoderSplattingCUDA.forward receives."""

    means: np.ndarray  # [P,3] world
    covariances: np.ndarray  # [P,3,3] world
    harmonics: np.ndarray  # [P,3,d_sh]
    opacities: np.ndarray  # [P]
    extrinsics: np.ndarray  # [4,4] camera-to-world (OpenCV axes, +z forward)
    intrinsics: np.ndarray  # [3,3] normalised
    near: float
    far: float
    image_shape: tuple  # (H, W)


def _random_rotations(rng, n):
    q = rng.standard_normal((n, 4)).astype(np.float64)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y),
        ],
        axis=1,
    ).reshape(n, 3, 3)
    return R


def small_se3(rng, rot_deg=3.0, trans=0.05):
    """A camera-to-world pose close to identity."""
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rot_deg) * rng.uniform(0.3, 1.0)
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * (Kx @ Kx)
    c2w = np.eye(4)
    c2w[:3, :3] = R
    c2w[:3, 3] = rng.uniform(-trans, trans, 3)
    return c2w.astype(np.float32)


def make_scene(P: int, H: int, W: int, sh_degree: int = 4, seed: int = SEED, fov_x_deg: float = 60.0,
               near: float = 1.0, far: float = 100.0, opacity_max: float = 1.0 / 3.0,
               behind_fraction: float = 0.0) -> Scene:
    rng = np.random.default_rng(seed)
    fx = 0.5 / np.tan(np.deg2rad(fov_x_deg) / 2)
    fy = fx * W / H  # square pixels: fx*W == fy*H
    Kn = np.array([[fx, 0, 0.5], [0, fy, 0.5], [0, 0, 1]], np.float32)
    c2w = small_se3(rng)

    # mean = unprojection of a uniform pixel (5 % margin outside the image) at uniform disparity
    u = rng.uniform(-0.05, 1.05, P)
    v = rng.uniform(-0.05, 1.05, P)
    disp = rng.uniform(1.0 / far, 1.0 / near, P)
    depth = 1.0 / disp
    if behind_fraction > 0:  # a few points behind / too close to the camera (cull path)
        m = rng.uniform(size=P) < behind_fraction
        depth = np.where(m, rng.uniform(-2.0, 0.2, P), depth)
    cam_pts = np.stack([(u - 0.5) / fx * depth, (v - 0.5) / fy * depth, depth], axis=1)
    means = cam_pts @ c2w[:3, :3].T.astype(np.float64) + c2w[:3, 3].astype(np.float64)

    # covariance = R diag(s^2) R^T, s = depth * 0.1 * pixel_size * (0.5 + 14.5 sigmoid(n))
    pixel = 1.0 / (fx * W) + 1.0 / (fy * H)
    mult = 0.5 + 14.5 / (1.0 + np.exp(-rng.standard_normal((P, 3))))
    s = np.abs(depth)[:, None] * 0.1 * pixel * mult
    R = _random_rotations(rng, P)
    cov = np.einsum("pij,pj,pkj->pik", R, s * s, R)

    opac = rng.uniform(0.01, opacity_max, P)
    d_sh = (sh_degree + 1) ** 2
    band = np.concatenate([np.full(2 * l + 1, 1.0 if l == 0 else 0.1 * 0.25**l) for l in range(sh_degree + 1)])
    harm = rng.standard_normal((P, 3, d_sh)) * band[None, None, :]

    return Scene(
        means=means.astype(np.float32),
        covariances=cov.astype(np.float32),
        harmonics=harm.astype(np.float32),
        opacities=opac.astype(np.float32),
        extrinsics=c2w,
        intrinsics=Kn,
        near=float(near),
        far=float(far),
        image_shape=(H, W),
    )


@dataclass
class RasterInputs:
    """Rasterizer-level inputs: exactly the arguments of the call at cuda_splatting.py:101-125."""

    means3D: np.ndarray  # [P,3]
    cov3D: np.ndarray  # [P,6]  (xx,xy,xz,yy,yz,zz)
    opacities: np.ndarray  # [P,1]
    shs: np.ndarray  # [P,K,3]
    viewmatrix: np.ndarray  # [4,4]
    projmatrix: np.ndarray  # [4,4]
    campos: np.ndarray  # [3]
    bg: np.ndarray  # [3]
    tanfovx: float
    tanfovy: float
    image_height: int
    image_width: int
    sh_degree: int


def to_raster_inputs(scene: Scene, bg=(0.0, 0.0, 0.0)) -> RasterInputs:
    """The camera / scaling maths of render_cuda (cuda_splatting.py:64-89) in float32 numpy."""
    f32 = np.float32
    H, W = scene.image_shape
    scale = f32(1.0) / f32(scene.near)
    extr = scene.extrinsics.astype(f32).copy()
    extr[:3, 3] = extr[:3, 3] * scale
    cov = scene.covariances.astype(f32) * (scale * scale)
    means = scene.means.astype(f32) * scale
    near, far = f32(scene.near) * scale, f32(scene.far) * scale
    Kn = scene.intrinsics.astype(f32)
    # get_fov (ggrt/geometry/projection.py:233-247) with cx = cy = 0.5: tan(fov/2) = 0.5 / f
    tanx = f32(0.5) / Kn[0, 0]
    tany = f32(0.5) / Kn[1, 1]
    proj = np.zeros((4, 4), f32)
    proj[0, 0] = 2 * near * Kn[0, 0]
    proj[1, 1] = 2 * near * Kn[1, 1]
    proj[0, 2] = 2 * Kn[0, 2] - 1
    proj[1, 2] = 2 * Kn[1, 2] - 1
    proj[3, 2] = 1
    proj[2, 2] = far / (far - near)
    proj[2, 3] = -(far * near) / (far - near)
    view = np.linalg.inv(extr.astype(np.float64)).astype(f32).T
    full = (view @ proj.T).astype(f32)
    iu = np.triu_indices(3)
    K = scene.harmonics.shape[-1]
    deg = int(round(np.sqrt(K))) - 1
    return RasterInputs(
        means3D=np.ascontiguousarray(means),
        cov3D=np.ascontiguousarray(cov[:, iu[0], iu[1]]),
        opacities=np.ascontiguousarray(scene.opacities.astype(f32)[:, None]),
        shs=np.ascontiguousarray(np.transpose(scene.harmonics.astype(f32), (0, 2, 1))),
        viewmatrix=np.ascontiguousarray(view),
        projmatrix=np.ascontiguousarray(full),
        campos=np.ascontiguousarray(extr[:3, 3]),
        bg=np.asarray(bg, f32),
        tanfovx=float(tanx),
        tanfovy=float(tany),
        image_height=H,
        image_width=W,
        sh_degree=deg,
    )


def image_gradient(H: int, W: int, seed: int = SEED, quadrant_only: bool = False) -> np.ndarray:
    """Upstream gradient dL/dimage ~ N(0,1)/(3HW) (MSE-like); optionally zero outside one
    quadrant, the sparsity of crop training (finetune_ggrt_stable.py:126-142)."""
    rng = np.random.default_rng(seed + 1)
    g = (rng.standard_normal((3, H, W)) / (3.0 * H * W)).astype(np.float32)
    if quadrant_only:
        m = np.zeros((H, W), bool)
        m[: H // 2, : W // 2] = True
        g = g * m[None]
    return g
