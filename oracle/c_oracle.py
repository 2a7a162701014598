"""ctypes front for oracle/raster_oracle.c (CPU restatement of the 3DGS tile rasterizer).

TEST INFRASTRUCTURE ONLY -- see the header of raster_oracle.c.  PARITY UNPINNED: no
reference binary / golden vectors exist for this path.

All arrays are numpy, C-contiguous; shapes follow the GGRt call site
(/root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py:101-125):
means [P,3], cov3d [P,6] (xx,xy,xz,yy,yz,zz), opacities [P], sh [P,K,3] or colors [P,3].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libraster_oracle.so"
_lib = None

TILE = 16


class _Cam(C.Structure):
    _fields_ = [
        ("P", C.c_int),
        ("deg", C.c_int),
        ("W", C.c_int),
        ("H", C.c_int),
        ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("view", C.c_float * 16),
        ("proj", C.c_float * 16),
        ("campos", C.c_float * 3),
        ("bg", C.c_float * 3),
    ]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = _HERE / "raster_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True, env={**os.environ, "CC": "/usr/bin/gcc"})
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.oracle_count_pairs.restype = C.c_uint64
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


@dataclass
class Camera:
    """Per-view settings, same meaning as GaussianRasterizationSettings (cuda_splatting.py:101-113)."""

    W: int
    H: int
    tanfovx: float
    tanfovy: float
    view: np.ndarray  # [4,4] as passed: transposed world->camera
    proj: np.ndarray  # [4,4] as passed: view @ projection^T
    campos: np.ndarray  # [3]
    bg: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    deg: int = 0

    def c(self, P: int) -> _Cam:
        cam = _Cam()
        cam.P, cam.deg, cam.W, cam.H = int(P), int(self.deg), int(self.W), int(self.H)
        cam.tanfovx, cam.tanfovy = float(self.tanfovx), float(self.tanfovy)
        cam.view[:] = np.asarray(self.view, np.float32).reshape(16).tolist()
        cam.proj[:] = np.asarray(self.proj, np.float32).reshape(16).tolist()
        cam.campos[:] = np.asarray(self.campos, np.float32).reshape(3).tolist()
        cam.bg[:] = np.asarray(self.bg, np.float32).reshape(3).tolist()
        return cam

    @property
    def grid(self):
        return (self.W + TILE - 1) // TILE, (self.H + TILE - 1) // TILE


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def preprocess(cam: Camera, means, cov3d, opacities, sh=None, colors=None) -> dict:
    L = lib()
    means, cov3d, opacities, sh, colors = _f32(means), _f32(cov3d), _f32(opacities).reshape(-1), _f32(sh), _f32(colors)
    P = means.shape[0]
    assert (sh is None) != (colors is None)
    if sh is not None:
        assert sh.shape == (P, (cam.deg + 1) ** 2, 3), sh.shape
    out = dict(
        depth=np.zeros(P, np.float32),
        radii=np.zeros(P, np.int32),
        xy=np.zeros((P, 2), np.float32),
        conic_opacity=np.zeros((P, 4), np.float32),
        rgb=np.zeros((P, 3), np.float32),
        clamped=np.zeros((P, 3), np.uint8),
        rect=np.zeros((P, 4), np.int32),
        tiles_touched=np.zeros(P, np.uint32),
    )
    cc = cam.c(P)
    L.oracle_preprocess(C.byref(cc), _p(means), _p(cov3d), _p(opacities), _p(sh), _p(colors), _p(out["depth"]),
                        _p(out["radii"]), _p(out["xy"]), _p(out["conic_opacity"]), _p(out["rgb"]), _p(out["clamped"]),
                        _p(out["rect"]), _p(out["tiles_touched"]))
    return out


def bin_tiles(cam: Camera, pre: dict) -> dict:
    L = lib()
    P = pre["depth"].shape[0]
    offsets = np.zeros(P, np.uint32)
    N = int(L.oracle_count_pairs(C.c_int(P), _p(pre["tiles_touched"]), _p(offsets)))
    gx, gy = cam.grid
    keys = np.zeros(max(N, 1), np.uint64)
    vals = np.zeros(max(N, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    L.oracle_bin(C.c_int(P), C.c_int(cam.W), C.c_int(cam.H), _p(pre["depth"]), _p(pre["radii"]), _p(pre["rect"]),
                 _p(offsets), C.c_uint64(N), _p(keys), _p(vals), _p(ranges))
    return dict(N=N, offsets=offsets, keys=keys[:N], point_list=vals[:N], ranges=ranges)


def render_forward(cam: Camera, pre: dict, binning: dict, aux=None) -> dict:
    """aux [P]: optional 4th blended channel (default: the view depth)."""
    L = lib()
    chan = pre["depth"] if aux is None else _f32(aux).reshape(-1)
    P = pre["depth"].shape[0]
    H, W = cam.H, cam.W
    out = dict(
        color=np.zeros((3, H, W), np.float32),
        depth=np.zeros((H, W), np.float32),
        final_T=np.zeros((H, W), np.float32),
        n_contrib=np.zeros((H, W), np.uint32),
        fragile=np.zeros((H, W), np.uint8),
        fragile_gaussian=np.zeros(P, np.uint8),
    )
    pl = binning["point_list"] if binning["N"] > 0 else np.zeros(1, np.uint32)
    cc = cam.c(P)
    L.oracle_render_forward(C.byref(cc), _p(binning["ranges"]), _p(pl), _p(pre["xy"]), _p(pre["conic_opacity"]),
                            _p(pre["rgb"]), _p(chan), _p(out["color"]), _p(out["depth"]), _p(out["final_T"]),
                            _p(out["n_contrib"]), _p(out["fragile"]), _p(out["fragile_gaussian"]))
    return out


def forward(cam: Camera, means, cov3d, opacities, sh=None, colors=None, aux=None) -> dict:
    """Whole forward: returns a dict with pre / bin / img sub-dicts plus color, depth, radii."""
    pre = preprocess(cam, means, cov3d, opacities, sh=sh, colors=colors)
    b = bin_tiles(cam, pre)
    img = render_forward(cam, pre, b, aux=aux)
    return dict(pre=pre, bin=b, img=img, color=img["color"], depth=img["depth"], radii=pre["radii"],
                aux=None if aux is None else _f32(aux).reshape(-1))


def backward(cam: Camera, means, cov3d, opacities, fwd: dict, dL_dcolor_img, sh=None, colors=None,
             dL_ddepth_img=None, want_camera: bool = False) -> dict:
    """Whole backward. Returns grads in the layout of upstream's autograd outputs.
    want_camera: also g["dcamera"] = dL/d(viewmatrix [16] | projmatrix [16] | campos [3]) as float64 [35]
    (the opt-in pose-gradient extension of BASELINE config 3)."""
    L = lib()
    means, cov3d, sh = _f32(means), _f32(cov3d), _f32(sh)
    P = means.shape[0]
    K = (cam.deg + 1) ** 2
    pre, b, img = fwd["pre"], fwd["bin"], fwd["img"]
    dimg = _f32(dL_dcolor_img)
    ddep = _f32(dL_ddepth_img)
    assert dimg.shape == (3, cam.H, cam.W)
    g = dict(
        dmean2D=np.zeros((P, 2), np.float32),
        dconic=np.zeros((P, 3), np.float32),
        dopacity=np.zeros(P, np.float32),
        dcolor=np.zeros((P, 3), np.float32),
        ddepth=np.zeros(P, np.float32) if ddep is not None else None,
    )
    pl = b["point_list"] if b["N"] > 0 else np.zeros(1, np.uint32)
    cc = cam.c(P)
    aux = fwd.get("aux")
    chan = pre["depth"] if aux is None else aux
    L.oracle_render_backward(C.byref(cc), _p(b["ranges"]), _p(pl), _p(pre["xy"]), _p(pre["conic_opacity"]),
                             _p(pre["rgb"]), _p(chan), _p(img["final_T"]), _p(img["n_contrib"]), _p(dimg),
                             _p(ddep), _p(g["dmean2D"]), _p(g["dconic"]), _p(g["dopacity"]), _p(g["dcolor"]),
                             _p(g["ddepth"]))
    g["dmeans3D"] = np.zeros((P, 3), np.float32)
    g["dcov3D"] = np.zeros((P, 6), np.float32)
    g["dsh"] = np.zeros((P, K, 3), np.float32) if sh is not None else None
    g["dcamera"] = np.zeros(35, np.float64) if want_camera else None
    L.oracle_preprocess_backward(C.byref(cc), _p(means), _p(cov3d), _p(sh), _p(pre["radii"]), _p(pre["clamped"]),
                                 _p(g["dmean2D"]), _p(g["dconic"]), _p(g["dcolor"]),
                                 _p(g["ddepth"]) if aux is None else None,  # a caller-supplied channel is a leaf
                                 _p(g["dmeans3D"]), _p(g["dcov3D"]), _p(g["dsh"]), _p(g["dcamera"]))
    g["daux"] = g["ddepth"] if aux is not None else None
    return g


def mark_visible(cam: Camera, means) -> np.ndarray:
    means = _f32(means)
    vis = np.zeros(means.shape[0], np.uint8)
    cc = cam.c(means.shape[0])
    lib().oracle_mark_visible(C.byref(cc), _p(means), _p(vis))
    return vis.astype(bool)
