"""One C2 (or other) fwd+bwd step for ncu: `ncu ... python tools/profile_step.py [workload] [warm steps]`."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ggrt_official_b200 import GaussianRasterizationSettings  # noqa: E402
from ggrt_official_b200 import rasterizer as R  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
ri, g_np = bench.make_inputs(workload, 0)
t = lambda a: torch.tensor(np.asarray(a), device=dev)
rs = GaussianRasterizationSettings(
    image_height=ri.image_height, image_width=ri.image_width, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(ri.bg),
    scale_modifier=1.0, viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix), sh_degree=ri.sh_degree,
    campos=t(ri.campos), prefiltered=False)
args = (t(ri.means3D), t(ri.shs), None, t(ri.opacities), t(ri.cov3D), rs)
g = t(g_np)
for _ in range(warm + 1):
    st = R.forward_raw(*args)
    R.backward_raw(st, g)
torch.cuda.synchronize()
print("N", st["N"])
