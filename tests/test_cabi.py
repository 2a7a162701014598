"""The C-ABI library loads, exports every symbol the header declares, and its host-side
helpers behave (no compute call: this runs without a GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from ggrt_official_b200 import _cabi

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "ggrt_raster.h").read_text()


def declared_functions():
    names = set(re.findall(r"\b(ggrt_(?:raster|adapter|camera)_\w+)\s*\(", HEADER))
    return sorted(names)


def test_header_cites_the_reference_interface():
    assert "cuda_splatting.py:6-9" in HEADER and ":101-125" in HEADER and "rasterize_gaussians" in HEADER


def test_every_declared_symbol_is_exported():
    lib = _cabi.lib()
    names = declared_functions()
    assert len(names) >= 12, names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ggrt_raster.h but not exported"
    assert set(_cabi.EXPORTS) == set(names)


def test_abi_version_and_layout():
    lib = _cabi.lib()
    assert lib.ggrt_raster_abi_version() == _cabi.ABI_VERSION
    P, H, W, N = 1000, 100, 75, 2500
    L = _cabi.layout(P, H, W, N)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    assert L.geom_bytes == lib.ggrt_raster_geom_bytes(P) >= P * (48 + 8 + 4 + 1 + 16 + 36) and L.geom_bytes >= L.geom_jac + 36 * P
    assert L.img_bytes == lib.ggrt_raster_image_bytes(H, W) >= 8 * H * W + 4 * (T + 1) + 2 * 4 * 32 * T
    assert L.bin_bytes == lib.ggrt_raster_binning_bytes(N) >= 12 * N
    offs = [L.geom_rec0, L.geom_rec1, L.geom_rec2, L.geom_rect, L.geom_tiles, L.geom_flags, L.geom_ranks, L.geom_jac]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    assert L.img_partials == L.img_counts + 4 * 32 * T and L.img_cursor > L.img_partials
    assert lib.ggrt_raster_binning_bytes(0) > 0  # never a zero-sized allocation


def test_argument_errors_are_reported_without_touching_cuda():
    lib = _cabi.lib()
    rc = lib.ggrt_raster_forward_prepare(None, None, 0, None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1
    assert b"settings" in lib.ggrt_raster_last_error()
    s = _cabi.Settings()
    s.image_height, s.image_width, s.tanfovx, s.tanfovy, s.sh_degree = 16, 16, 1.0, 1.0, 7
    dummy = C.c_void_p(16)
    s.viewmatrix = s.projmatrix = s.campos = s.bg = 16
    rc = lib.ggrt_raster_forward_prepare(C.byref(s), None, 1, dummy, dummy, dummy, dummy, None, None, dummy, dummy, dummy,
                                         None, None)
    assert rc == -3 and b"sh_degree" in lib.ggrt_raster_last_error()
    with pytest.raises(RuntimeError, match="sh_degree"):
        _cabi.check(rc, "forward_prepare")
    assert lib.ggrt_raster_stage_name(6) == b"render_backward"


def test_exchange_entry_points_validate_arguments():
    lib = _cabi.lib()
    ptrs = (C.c_void_p * 1)(16)
    dummy = C.c_void_p(16)
    assert lib.ggrt_raster_sh_gradient_merge(8, 4, None, dummy, 0, ptrs, ptrs, dummy, None) == -1
    assert b"num_views" in lib.ggrt_raster_last_error()
    assert lib.ggrt_raster_sh_gradient_merge(8, 4, None, dummy, _cabi.MAX_MERGE_VIEWS + 1, ptrs, ptrs, dummy, None) == -1
    assert lib.ggrt_raster_sh_gradient_merge(8, 5, None, dummy, 1, ptrs, ptrs, dummy, None) == -1
    null = (C.c_void_p * 1)(None)
    assert lib.ggrt_raster_sh_gradient_merge(8, 4, None, dummy, 1, null, ptrs, dummy, None) == -1
    assert b"view 0" in lib.ggrt_raster_last_error()
    assert lib.ggrt_raster_nvls_allreduce_f32(dummy, 6, 0, 2, None) == -1  # not a multiple of 4
    assert lib.ggrt_raster_nvls_allreduce_f32(dummy, 8, 2, 2, None) == -1  # rank outside the world
    assert lib.ggrt_raster_nvls_allreduce_f32(None, 8, 0, 2, None) == -1


def test_upstream_structure_baseline_builds_and_exports():
    """The GPU baseline (measurement aid, never imported by the product) compiles for sm_100a and links the product."""
    from baseline import build as ub

    lib = C.CDLL(str(ub.build()))
    for name in ("upstream_forward", "upstream_backward", "upstream_last_error"):
        assert hasattr(lib, name)


def test_ctypes_signatures_have_the_arity_the_header_declares():
    """Every prototype of include/ggrt_raster.h against the argtypes the ctypes binding registers."""
    lib = _cabi.lib()
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    protos = re.findall(r"\b(ggrt_(?:raster|adapter|camera)_\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 16
    checked = 0
    for name, params in protos:
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        fn = getattr(lib, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, (name, len(fn.argtypes), n)
            checked += 1
    assert checked >= 12


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct in include/ggrt_raster.h, as gcc lays them out, against the ctypes mirrors in
    _cabi.py (a field added on one side only would shift everything behind it silently)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"GgrtRasterSettings": _cabi.Settings, "GgrtRasterInputLayout": _cabi.InputLayout,
               "GgrtRasterGradSinks": _cabi.GradSinks, "GgrtAdapterParams": _cabi.AdapterParams,
               "GgrtRasterLayout": _cabi.Layout}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "ggrt_raster.h"}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, val = line.split()
        cls = structs[cname]
        expect = C.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(val) == expect, (cname, field, int(val), expect)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())
    # the header's constants the binding repeats
    assert f"#define GGRT_RASTER_ABI_VERSION {_cabi.ABI_VERSION}" in HEADER
    assert f"#define GGRT_RASTER_MAX_MERGE_VIEWS {_cabi.MAX_MERGE_VIEWS}" in HEADER
    assert f"#define GGRT_CAMERA_FLOATS {_cabi.CAMERA_FLOATS}" in HEADER and f"#define GGRT_STAGE_COUNT {_cabi.STAGE_COUNT}" in HEADER
