"""Differentiable PyTorch restatement of the rasterizer.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see raster_oracle.c).  This second,
independent restatement exists to validate the C oracle: its gradients come from
autograd of the forward, not from the hand-derived A.4/A.5 formulas.  Dense
[pixels x Gaussians] evaluation -- small cases only.

Call-site semantics: /root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py:101-125.
"""
from __future__ import annotations

import math

import torch

TILE = 16
NEAR_CULL = 0.2
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]
C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
      -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761]


def sh_basis(deg: int, d: torch.Tensor) -> torch.Tensor:
    """[P,3] unit directions -> [P,K] real SH basis, 3DGS / PlenOctrees sign convention (A.1)."""
    x, y, z = d.unbind(-1)
    one = torch.ones_like(x)
    b = [C0 * one]
    if deg > 0:
        b += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg > 2:
        b += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
              C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
              C3[6] * x * (xx - 3 * yy)]
    if deg > 3:
        b += [C4[0] * xy * (xx - yy), C4[1] * yz * (3 * xx - yy), C4[2] * xy * (7 * zz - 1),
              C4[3] * yz * (7 * zz - 3), C4[4] * (zz * (35 * zz - 30) + 3), C4[5] * xz * (7 * zz - 3),
              C4[6] * (xx - yy) * (7 * zz - 1), C4[7] * xz * (xx - 3 * yy),
              C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(b, dim=-1)


def rasterize(means, cov6, opac, *, view, proj, campos, bg, tanfovx, tanfovy, H, W, deg, sh=None, colors=None):
    """Returns (color [3,H,W], depth [H,W], aux).  All tensor args share one dtype (use float64)."""
    dt = means.dtype
    P = means.shape[0]
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    ones = torch.ones(P, 1, dtype=dt)
    ph = torch.cat([means, ones], dim=1)
    t = ph @ view  # [P,4]
    hom = ph @ proj
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    tz = t[:, 2]
    fx = W / (2.0 * tanfovx)
    fy = H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tzs = torch.where(tz > NEAR_CULL, tz, torch.ones_like(tz))  # keep culled rows finite
    cx = torch.clamp(t[:, 0] / tzs, -limx, limx) * tzs
    cy = torch.clamp(t[:, 1] / tzs, -limy, limy) * tzs
    zero = torch.zeros_like(tz)
    J = torch.stack(
        [torch.stack([fx / tzs, zero, -(fx * cx) / (tzs * tzs)], -1),
         torch.stack([zero, fy / tzs, -(fy * cy) / (tzs * tzs)], -1)], dim=1)  # [P,2,3]
    Rw = view[:3, :3].T  # world -> camera rotation
    Tm = J @ Rw  # [P,2,3]
    S = torch.stack(
        [torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2]], -1),
         torch.stack([cov6[:, 1], cov6[:, 3], cov6[:, 4]], -1),
         torch.stack([cov6[:, 2], cov6[:, 4], cov6[:, 5]], -1)], dim=1)
    c2 = Tm @ S @ Tm.transpose(1, 2)
    a = c2[:, 0, 0] + 0.3
    b = c2[:, 0, 1]
    c = c2[:, 1, 1] + 0.3
    det = a * c - b * b
    valid = (tz > NEAR_CULL) & (det != 0)
    dets = torch.where(valid, det, torch.ones_like(det))
    cA, cB, cC = c / dets, -b / dets, a / dets
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.clamp(lam, min=0.0))).detach()
    pix = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], -1)
    pd = pix.detach()

    def tl(v, g):
        return torch.clamp(torch.trunc(torch.clamp(v / TILE, -1, 65536)), 0, g).long()

    x0, y0 = tl(pd[:, 0] - radius, gx), tl(pd[:, 1] - radius, gy)
    x1, y1 = tl(pd[:, 0] + radius + (TILE - 1), gx), tl(pd[:, 1] + radius + (TILE - 1), gy)
    valid = valid & ((x1 - x0) * (y1 - y0) > 0)

    if colors is not None:
        rgb = colors
        clamped = torch.zeros(P, 3, dtype=torch.bool)
    else:
        d = means - campos[None]
        d = d / d.norm(dim=-1, keepdim=True)
        rgb = torch.einsum("pk,pkc->pc", sh_basis(deg, d), sh) + 0.5
        clamped = rgb < 0
        rgb = torch.clamp(rgb, min=0.0)

    # global (depth, index) order == per-tile order restricted to each tile's list (A.2)
    order = torch.argsort(tz.detach(), stable=True)
    order = order[valid[order]]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pxf, pyf = xs.reshape(-1).to(dt), ys.reshape(-1).to(dt)  # [HW]
    ptx, pty = (xs.reshape(-1) // TILE), (ys.reshape(-1) // TILE)
    o = order
    in_tile = ((ptx[:, None] >= x0[o][None]) & (ptx[:, None] < x1[o][None]) &
               (pty[:, None] >= y0[o][None]) & (pty[:, None] < y1[o][None]))  # [HW,G]
    dx = pix[o, 0][None] - pxf[:, None]
    dy = pix[o, 1][None] - pyf[:, None]
    power = -0.5 * (cA[o][None] * dx * dx + cC[o][None] * dy * dy) - cB[o][None] * dx * dy
    alpha = torch.clamp(opac.reshape(-1)[o][None] * torch.exp(torch.clamp(power, max=0.0)), max=0.99)
    live = in_tile & (power <= 0) & (alpha >= 1.0 / 255.0)
    alpha = torch.where(live, alpha, torch.zeros_like(alpha))
    Tn = torch.cumprod(1.0 - alpha, dim=1)  # transmittance AFTER each Gaussian
    # "if T*(1-alpha) < 1e-4: break" -- that Gaussian and everything behind it is dropped
    dead = torch.cumsum(((Tn < 1e-4) & live).to(torch.int32), dim=1) > 0
    alpha = torch.where(dead, torch.zeros_like(alpha), alpha)
    Tn = torch.cumprod(1.0 - alpha, dim=1)
    Tb = torch.cat([torch.ones(H * W, 1, dtype=dt), Tn[:, :-1]], dim=1)  # transmittance BEFORE
    wgt = alpha * Tb
    color = wgt @ rgb[o] + Tn[:, -1:] * bg[None] if o.numel() else bg[None].expand(H * W, 3)
    depth = wgt @ tz[o] if o.numel() else torch.zeros(H * W, dtype=dt)
    final_T = Tn[:, -1] if o.numel() else torch.ones(H * W, dtype=dt)
    aux = dict(radii=torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32), valid=valid,
               rect=torch.stack([x0, y0, x1, y1], -1), pix=pix, conic=torch.stack([cA, cB, cC], -1), rgb=rgb,
               clamped=clamped, depth=tz, final_T=final_T.reshape(H, W), order=order,
               n_live=(live & ~dead).sum(1).reshape(H, W))
    return color.T.reshape(3, H, W), depth.reshape(H, W), aux
