"""Single-GPU timing of the exchange kernels with virtual views (C2 shape): the per-Gaussian backward in compact mode
and sh_gradient_merge at V = 1, 2, 4, 8 local views.  CUDA events, L2 flushed.  One JSON line."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ggrt_official_b200 import GaussianRasterizationSettings, _cabi  # noqa: E402
from ggrt_official_b200 import rasterizer as R  # noqa: E402

dev = torch.device("cuda:0")
P, H, W, _ = bench.WORKLOADS["c2"]
ri, g_np = bench.make_inputs("c2", 0)
t = lambda a: torch.tensor(np.asarray(a), device=dev)
means, cov, opac, shs, g = t(ri.means3D), t(ri.cov3D), t(ri.opacities), t(ri.shs), t(g_np)
rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(ri.bg),
                                   scale_modifier=1.0, viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix),
                                   sh_degree=ri.sh_degree, campos=t(ri.campos), prefiltered=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = R.forward_raw(means, shs, None, opac, cov, rs)
full = R.backward_raw(st, g)
comp = R.backward_raw(st, g, compact=True)
out = {}


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ms = bench.time_loop(fn, n, dev, before=flush.zero_)
    return sum(ms) / n


_cabi.profile_enable(True)
acc = {"full": 0.0, "compact": 0.0}
for _ in range(10):
    flush.zero_()
    R.backward_raw(st, g)
    acc["full"] += _cabi.profile_read()["preprocess_backward"] / 10
    flush.zero_()
    R.backward_raw(st, g, compact=True)
    acc["compact"] += _cabi.profile_read()["preprocess_backward"] / 10
_cabi.profile_enable(False)
out["preprocess_backward_ms"] = acc
dsh = torch.empty_like(full["dsh"])
for V in (1, 2, 4, 8):
    drgb = [comp["dcolors"] * (1.0 + 0.1 * v) for v in range(V)]
    cams = [rs.campos + 0.01 * v for v in range(V)]
    out[f"merge_V{V}_ms"] = timed(lambda: R.sh_gradient_merge(means, 4, drgb, cams, out=dsh))
ref = sum(full["dsh"] * 0 + 0 for _ in range(1))  # keep the allocator warm
print(json.dumps(out))
