"""GPU diagnostics: per-stage comparison with the oracle on several cases + per-stage timings.
Writes gpurun_out/diag.json.  Run on the B200 box: python tools/gpu_diag.py"""
import json
import sys
import time
import traceback
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ggrt_official_b200 import rasterizer as R  # noqa: E402
from ggrt_official_b200.synthetic import image_gradient, make_scene, to_raster_inputs  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from tests import gpu_util as G  # noqa: E402
from tests.helpers import small_case  # noqa: E402

out = {}


def case(name, ri, grad_seed=0, timing=False):
    rec = {}
    try:
        H, W = ri.image_height, ri.image_width
        st = G.run_cuda_forward(ri, debug=True)
        t0 = time.time()
        cam, f = G.oracle_forward(ri)
        rec["oracle_fwd_s"] = time.time() - t0
        rec["forward"] = G.compare_forward(st, f)
        g = np.random.default_rng(grad_seed).standard_normal((3, H, W)).astype(np.float32) / (3 * H * W)
        got = R.backward_raw(st, torch.tensor(g, device="cuda:0"))
        torch.cuda.synchronize()
        t0 = time.time()
        ref = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs)
        rec["oracle_bwd_s"] = time.time() - t0
        rec["backward"] = G.grad_errors(got, ref)
        if timing:
            rec["timing_ms"] = time_stages(ri)
    except Exception:
        rec["error"] = traceback.format_exc()
    out[name] = rec
    print(name, json.dumps(rec, indent=1, default=str))


def time_stages(ri, iters=20):
    dev = "cuda:0"
    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    rs = G.settings_from(ri, dev)
    args = (t(ri.means3D), t(ri.shs), None, t(ri.opacities), t(ri.cov3D), rs)
    g = t(image_gradient(ri.image_height, ri.image_width))
    for _ in range(3):
        st = R.forward_raw(*args)
        R.backward_raw(st, g)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fw, bw = [], []
    for _ in range(iters):
        ev[0].record()
        st = R.forward_raw(*args)
        ev[1].record()
        R.backward_raw(st, g)
        ev[2].record()
        torch.cuda.synchronize()
        fw.append(ev[0].elapsed_time(ev[1]))
        bw.append(ev[1].elapsed_time(ev[2]))
    return dict(forward_med=float(np.median(fw)), backward_med=float(np.median(bw)), forward_min=float(min(fw)),
                backward_min=float(min(bw)), N=st["N"], max_tile_pairs=st["max_tile_pairs"])


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    _, ri = small_case(2000, 64, 80, 4, seed=1)
    case("small_deg4", ri)
    _, ri = small_case(3000, 100, 75, 4, bg=(0.2, 0.5, 0.7), seed=2, cov_scale=9.0)
    case("ragged_bg", ri)
    _, ri = small_case(500, 16, 16, 1, seed=6, cov_scale=100.0)
    case("single_tile", ri)
    case("C1_10k_256", to_raster_inputs(make_scene(10_000, 256, 256, sh_degree=4)), timing=True)
    case("C2_300k_1008x756", to_raster_inputs(make_scene(300_000, 756, 1008, sh_degree=4)), timing=True)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "diag.json").write_text(json.dumps(out, indent=1, default=str))
