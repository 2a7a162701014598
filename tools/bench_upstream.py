"""Times the "upstream-structure" GPU baseline (baseline/upstream_structure.cu: prefix sum + host read of N, global
64-bit radix sort, 16x16-thread render kernels, per-thread global atomics in the backward) next to the product path on
the same B200 and the same inputs, and checks that both produce the same image and gradients.
NOT the reference binary (its source is not available offline) -- see the header of upstream_structure.cu.

  python tools/bench_upstream.py [--workload c2] [--steps 30]        -> one JSON line
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from baseline import build as ub  # noqa: E402
from ggrt_official_b200 import GaussianRasterizationSettings, _cabi  # noqa: E402
from ggrt_official_b200 import rasterizer as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    P, H, W, desc = bench.WORKLOADS[args.workload]
    ri, g_np = bench.make_inputs(args.workload, 0)
    K = (bench.SH_DEGREE + 1) ** 2
    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    means, cov, opac, shs, grad_img = t(ri.means3D), t(ri.cov3D), t(ri.opacities).reshape(-1).contiguous(), t(ri.shs), t(g_np)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(ri.bg), scale_modifier=1.0,
        viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix), sh_degree=ri.sh_degree, campos=t(ri.campos),
        prefiltered=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(step, n):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / n

    # ---- product ----
    def ours():
        st = R.forward_raw(means, shs, None, opac, cov, rs)
        return st, R.backward_raw(st, grad_img)

    st, go = ours()
    torch.cuda.synchronize()
    ms_ours = timed(ours, args.steps)

    # ---- upstream structure ----
    L = C.CDLL(str(ub.build()))
    L.upstream_last_error.restype = C.c_char_p
    vp = C.c_void_p
    L.upstream_forward.argtypes = [C.POINTER(_cabi.Settings), C.c_int] + [vp] * 7 + [C.POINTER(C.c_longlong)]
    L.upstream_backward.argtypes = [C.POINTER(_cabi.Settings), C.c_int] + [vp] * 10
    call = st["call"]
    s = call.settings
    f32 = dict(dtype=torch.float32, device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    color, depth = torch.empty((3, H, W), **f32), torch.empty((H, W), **f32)
    gb = dict(dmeans2D=torch.empty((P, 3), **f32), dopacity=torch.empty((P, 1), **f32), dmeans3D=torch.empty((P, 3), **f32),
              dcov3D=torch.empty((P, 6), **f32), dsh=torch.empty((P, K, 3), **f32))
    p = lambda x: C.c_void_p(x.data_ptr())
    n_out = C.c_longlong(0)

    def check(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {L.upstream_last_error().decode()}")

    def upstream():
        check(L.upstream_forward(C.byref(s), P, p(call.means3D), p(call.cov3D), p(call.opacities), p(call.sh), p(radii),
                                 p(color), p(depth), C.byref(n_out)), "upstream_forward")
        check(L.upstream_backward(C.byref(s), P, p(call.means3D), p(call.cov3D), p(call.sh), p(radii), p(grad_img),
                                  p(gb["dmeans2D"]), p(gb["dopacity"]), p(gb["dmeans3D"]), p(gb["dcov3D"]), p(gb["dsh"])),
              "upstream_backward")

    with torch.cuda.stream(torch.cuda.default_stream(dev)):  # upstream launches on the legacy default stream
        upstream()
        torch.cuda.synchronize()
        ms_up = timed(upstream, args.steps)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    diffs = {k: rel(gb[k], go[k]) for k in gb}
    line = {
        "what": "upstream-structure GPU baseline vs product, fwd+bwd, same B200, same inputs (NOT the reference binary)",
        "workload": desc, "P": P, "N_pairs": int(n_out.value), "N_pairs_product": int(st["N"]),
        "upstream_structure_ms": round(ms_up, 4), "product_ms": round(ms_ours, 4), "speedup": round(ms_up / ms_ours, 2),
        "upstream_structure_fps": round(1e3 / ms_up, 1), "product_fps": round(1e3 / ms_ours, 1),
        "radii_equal": bool(torch.equal(radii, st["radii"])),
        "color_max_abs_diff": float((color - st["color"]).abs().max()),
        "depth_max_rel_diff": rel(depth, st["depth"]), "grad_max_rel_diff": diffs,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
