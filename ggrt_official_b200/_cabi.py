"""ctypes binding of libggrt_raster.so (include/ggrt_raster.h).

There is no CPU fallback: if the library is missing or a CUDA device is not present the
calls raise.  The library is built in-tree by ggrt_official_b200.build (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

ABI_VERSION = 11
_lib = None


class Settings(C.Structure):
    """struct GgrtRasterSettings"""

    _fields_ = [
        ("image_height", C.c_int32),
        ("image_width", C.c_int32),
        ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("scale_modifier", C.c_float),
        ("sh_degree", C.c_int32),
        ("prefiltered", C.c_int32),
        ("debug", C.c_int32),
        ("viewmatrix", C.c_void_p),
        ("projmatrix", C.c_void_p),
        ("campos", C.c_void_p),
        ("bg", C.c_void_p),
        ("device_params", C.c_void_p),
        ("aux_mode", C.c_int32),
        ("reserved", C.c_int32),
        ("zero_scratch", C.c_void_p),
    ]


class InputLayout(C.Structure):
    """struct GgrtRasterInputLayout"""

    _fields_ = [("scene_scale", C.c_float), ("cov_full3x3", C.c_int32), ("sh_channel_major", C.c_int32)]


class GradSinks(C.Structure):
    """struct GgrtRasterGradSinks"""

    _fields_ = [("count", C.c_int32), ("multimem", C.c_int32), ("ptr", C.c_void_p * 16),
                ("epoch", C.c_void_p), ("done_counter", C.c_void_p), ("parity_stride", C.c_int64),
                ("arrive_count", C.c_int32), ("reserved", C.c_int32), ("arrive", C.c_void_p * 16)]


class AdapterParams(C.Structure):
    """struct GgrtAdapterParams"""

    _fields_ = [("num_views", C.c_int32), ("rays_per_view", C.c_int32), ("samples_per_ray", C.c_int32),
                ("sh_degree", C.c_int32), ("image_height", C.c_int32), ("image_width", C.c_int32),
                ("scale_min", C.c_float), ("scale_max", C.c_float), ("eps", C.c_float)]


class Layout(C.Structure):
    """struct GgrtRasterLayout"""

    _fields_ = [(n, C.c_size_t) for n in (
        "geom_rec0", "geom_rec1", "geom_rec2", "geom_rect", "geom_tiles", "geom_flags", "geom_ranks", "geom_jac", "geom_bytes",
        "img_counts", "img_partials", "img_cursor", "img_starts", "img_header", "img_final_T", "img_ncontrib", "img_bytes",
        "bin_keys", "bin_points", "bin_masks", "bin_bytes")]


EXPORTS = (
    "ggrt_raster_abi_version",
    "ggrt_raster_last_error",
    "ggrt_raster_layout",
    "ggrt_raster_geom_bytes",
    "ggrt_raster_image_bytes",
    "ggrt_raster_binning_bytes",
    "ggrt_raster_forward_prepare",
    "ggrt_raster_forward_render",
    "ggrt_raster_join",
    "ggrt_raster_backward",
    "ggrt_raster_sh_gradient_merge",
    "ggrt_raster_sh_gradient_merge_signalled",
    "ggrt_raster_nvls_allreduce_f32",
    "ggrt_raster_nvls_allreduce_signalled",
    "ggrt_raster_nvls_barrier",
    "ggrt_adapter_forward",
    "ggrt_adapter_backward",
    "ggrt_camera_setup",
    "ggrt_raster_mark_visible",
    "ggrt_raster_profile_enable",
    "ggrt_raster_profile_read",
    "ggrt_raster_stage_name",
)
STAGE_COUNT = 8
MAX_MERGE_VIEWS = 16
CAMERA_FLOATS = 48


def library_path() -> Path:
    return _build.LIB


def lib():
    """Load (building if the sources are newer) the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    import os

    override = os.environ.get("GGRT_RASTER_LIB")  # tuning experiments: a library built with other -D knobs
    path = Path(override) if override else _build.build()
    L = C.CDLL(str(path))
    vp, i32, i64, u32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_size_t
    L.ggrt_raster_abi_version.restype = C.c_int
    L.ggrt_raster_last_error.restype = C.c_char_p
    L.ggrt_raster_layout.argtypes = [i32, i32, i32, i64, C.POINTER(Layout)]
    L.ggrt_raster_geom_bytes.argtypes = [i32]
    L.ggrt_raster_geom_bytes.restype = sz
    L.ggrt_raster_image_bytes.argtypes = [i32, i32]
    L.ggrt_raster_image_bytes.restype = sz
    L.ggrt_raster_binning_bytes.argtypes = [i64]
    L.ggrt_raster_binning_bytes.restype = sz
    L.ggrt_raster_forward_prepare.argtypes = [C.POINTER(Settings), C.POINTER(InputLayout), i32] + [vp] * 11
    L.ggrt_raster_forward_render.argtypes = [C.POINTER(Settings), i32, i64, u32, i32, vp, vp, vp, vp, vp, vp]
    L.ggrt_raster_join.argtypes = [vp]
    L.ggrt_raster_backward.argtypes = ([C.POINTER(Settings), C.POINTER(InputLayout), i32, i64] + [vp] * 18
                                       + [C.POINTER(GradSinks), vp])
    L.ggrt_raster_sh_gradient_merge.argtypes = [i32, i32, C.POINTER(InputLayout), vp, i32, C.POINTER(vp), C.POINTER(vp),
                                                vp, vp]
    L.ggrt_raster_sh_gradient_merge_signalled.argtypes = [i32, i32, C.POINTER(InputLayout), vp, i32, vp, i64, i64, vp, vp,
                                                          vp, vp]
    L.ggrt_raster_nvls_allreduce_f32.argtypes = [vp, i64, i32, i32, vp]
    L.ggrt_raster_nvls_allreduce_signalled.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp]
    L.ggrt_raster_nvls_barrier.argtypes = [vp, vp, u32, vp]
    L.ggrt_adapter_forward.argtypes = [C.POINTER(AdapterParams)] + [vp] * 12
    L.ggrt_adapter_backward.argtypes = [C.POINTER(AdapterParams)] + [vp] * 13
    L.ggrt_camera_setup.argtypes = [i32, vp, vp, vp, vp, i32, vp, vp]
    L.ggrt_raster_mark_visible.argtypes = [i32, vp, vp, vp, vp]
    L.ggrt_raster_profile_enable.argtypes = [i32]
    L.ggrt_raster_profile_read.argtypes = [C.POINTER(C.c_float)]
    L.ggrt_raster_stage_name.argtypes = [i32]
    L.ggrt_raster_stage_name.restype = C.c_char_p
    if L.ggrt_raster_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libggrt_raster ABI {L.ggrt_raster_abi_version()} != expected {ABI_VERSION}")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ggrt_raster_last_error().decode(errors="replace")
        raise RuntimeError(f"libggrt_raster {what} failed (code {rc}): {msg}")


def layout(P: int, H: int, W: int, N: int) -> Layout:
    out = Layout()
    check(lib().ggrt_raster_layout(P, H, W, N, C.byref(out)), "layout")
    return out


def profile_enable(on: bool) -> None:
    check(lib().ggrt_raster_profile_enable(int(bool(on))), "profile_enable")


def profile_read() -> dict:
    """Per-kernel milliseconds of the most recent forward/backward on this thread."""
    buf = (C.c_float * STAGE_COUNT)()
    check(lib().ggrt_raster_profile_read(buf), "profile_read")
    return {lib().ggrt_raster_stage_name(i).decode(): float(buf[i]) for i in range(STAGE_COUNT)}
