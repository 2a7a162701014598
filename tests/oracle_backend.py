"""Test-only CPU backend: patches the two C-ABI entry points of the Python front
(`forward_raw` / `backward_raw`) with the CPU oracle, so the host-side logic (argument
validation, contiguity handling, autograd glue, view sharding) and the UNMODIFIED reference
caller can be exercised without a GPU.  The product never uses this: it lives under tests/.
"""
from __future__ import annotations

import contextlib
from types import SimpleNamespace

import numpy as np
import torch

from ggrt_official_b200 import rasterizer as R
from oracle import c_oracle as co

CALLS = []  # every forward call is recorded here (inputs as passed at the boundary)


def _np(t):
    return None if t is None else t.detach().cpu().contiguous().numpy()


def _camera(rs, deg):
    return co.Camera(W=int(rs.image_width), H=int(rs.image_height), tanfovx=float(rs.tanfovx),
                     tanfovy=float(rs.tanfovy), view=_np(rs.viewmatrix), proj=_np(rs.projmatrix),
                     campos=_np(rs.campos), bg=_np(rs.bg), deg=deg)


def fake_forward_raw(means3D, sh, colors_precomp, opacities, cov3D_precomp, rs, aux=None, layout=None,
                     workspace=None, check="sync", prezero_scratch=False):
    CALLS.append(dict(means3D=means3D, sh=sh, colors_precomp=colors_precomp, opacities=opacities,
                      cov3D_precomp=cov3D_precomp, settings=rs, aux=aux, layout=layout))
    deg = int(rs.sh_degree)
    cam = _camera(rs, deg)
    lay = layout or {}
    scale = np.float32(lay.get("scene_scale", 1.0))
    shn = _np(sh)
    if shn is not None:
        if lay.get("sh_channel_major"):
            shn = np.transpose(shn, (0, 2, 1))
        shn = np.ascontiguousarray(shn[:, : (deg + 1) ** 2, :])
    covn = _np(cov3D_precomp)
    if lay.get("cov_full3x3"):
        iu = np.triu_indices(3)
        covn = covn[:, iu[0], iu[1]]
    covn = np.ascontiguousarray(covn * (scale * scale))
    inp = dict(means=np.ascontiguousarray(_np(means3D) * scale), cov=covn, opac=_np(opacities).reshape(-1), sh=shn,
               colors=_np(colors_precomp), layout=dict(lay))
    auxn = None if aux is None else _np(aux).reshape(-1)
    f = co.forward(cam, inp["means"], inp["cov"], inp["opac"], sh=inp["sh"], colors=inp["colors"], aux=auxn)
    call = SimpleNamespace(means3D=means3D, sh=sh, colors=colors_precomp, opacities=opacities, cov3D=cov3D_precomp,
                           P=inp["means"].shape[0], aux=aux, H=int(rs.image_height), W=int(rs.image_width),
                           device=means3D.device, campos=rs.campos)
    return dict(call=call, color=torch.from_numpy(f["color"].copy()), depth=torch.from_numpy(f["depth"].copy()),
                radii=torch.from_numpy(f["radii"].copy()), geom=None, img=None, binning=None, N=f["bin"]["N"],
                capacity=f["bin"]["N"], max_tile_pairs=0, workspace=None, scratch=None, _oracle=(cam, inp, f))


def fake_backward_raw(state, grad_color, out=None, grad_aux=None, want_camera=False):
    assert not want_camera, "camera gradients are only implemented on the CUDA path"
    cam, inp, f = state["_oracle"]
    g = co.backward(cam, inp["means"], inp["cov"], inp["opac"], f, _np(grad_color), sh=inp["sh"], colors=inp["colors"],
                    dL_ddepth_img=_np(grad_aux))
    P = inp["means"].shape[0]
    t = torch.from_numpy
    d2 = np.zeros((P, 3), np.float32)
    d2[:, :2] = g["dmean2D"]
    lay = inp["layout"]
    scale = np.float32(lay.get("scene_scale", 1.0))
    g["dmeans3D"] = np.ascontiguousarray(g["dmeans3D"] * scale)
    g["dcov3D"] = np.ascontiguousarray(g["dcov3D"] * (scale * scale))
    if lay.get("cov_full3x3"):
        full = np.zeros((P, 3, 3), np.float32)
        iu = np.triu_indices(3)
        full[:, iu[0], iu[1]] = g["dcov3D"]
        g["dcov3D"] = full
    if lay.get("sh_channel_major") and g["dsh"] is not None:
        g["dsh"] = np.ascontiguousarray(np.transpose(g["dsh"], (0, 2, 1)))
    return dict(dmeans2D=t(d2), dopacity=t(g["dopacity"].reshape(P, 1).copy()), dmeans3D=t(g["dmeans3D"]),
                dcov3D=t(g["dcov3D"]), dsh=None if g["dsh"] is None else t(g["dsh"]),
                dcolors=None if inp["colors"] is None else t(g["dcolor"]),
                daux=None if g.get("daux") is None else t(g["daux"]))


@contextlib.contextmanager
def installed():
    """Route the rasterizer front through the CPU oracle for the duration of the block."""
    orig = (R.forward_raw, R.backward_raw, R._RasterizeGaussians.forward)
    R.forward_raw, R.backward_raw = fake_forward_raw, fake_backward_raw
    orig_fwd = R._RasterizeGaussians.forward

    def fwd(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
            aux=None, layout=None, viewmatrix=None, projmatrix=None, campos=None):
        out = orig_fwd(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                       raster_settings, aux, layout, viewmatrix, projmatrix, campos)
        ctx.state["_oracle"] = CALLS_STATE.pop()
        return out

    CALLS_STATE = []
    real_fake = fake_forward_raw

    def recording_forward(*a, **k):
        st = real_fake(*a, **k)
        CALLS_STATE.append(st["_oracle"])
        return st

    R.forward_raw = recording_forward
    R._RasterizeGaussians.forward = staticmethod(fwd)
    del CALLS[:]
    try:
        yield CALLS
    finally:
        R.forward_raw, R.backward_raw = orig[0], orig[1]
        R._RasterizeGaussians.forward = staticmethod(orig[2])


def fake_backward_compact(state, grad_color, out=None, compact=False, **kw):
    """`backward_raw(..., compact=True)` on the oracle: dcolors = masked colour gradient, no dsh."""
    g = fake_backward_raw(state, grad_color, **kw)
    if compact:
        cam, inp, f = state["_oracle"]
        pre = f["pre"]
        d = co.backward(cam, inp["means"], inp["cov"], inp["opac"], f, _np(grad_color), sh=inp["sh"])["dcolor"]
        d = d * (pre["clamped"] == 0) * (pre["radii"] > 0)[:, None]
        g["dcolors"], g["dsh"] = torch.from_numpy(d.astype(np.float32)), None
    if out:
        for k, dst in out.items():
            if g.get(k) is not None:
                dst.copy_(g[k].reshape(dst.shape))
                g[k] = dst
    return g


def fake_sh_gradient_merge(means3D, sh_degree, drgb_views, campos_views, out=None, layout=None):
    """`sh_gradient_merge` restated with torch on the CPU (float64 accumulate)."""
    from oracle.torch_ref import sh_basis

    scale = float((layout or {}).get("scene_scale", 1.0))
    m = means3D.detach().double() * scale
    acc = 0
    for d, cpos in zip(drgb_views, campos_views):
        dirs = m - cpos.detach().double().reshape(1, 3)
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        acc = acc + sh_basis(int(sh_degree), dirs)[:, :, None] * d.detach().double()[:, None, :]
    res = acc.float()
    if out is not None:
        out.copy_(res)
        return out
    return res
