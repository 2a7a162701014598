"""Small fwd+bwd for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ggrt_official_b200 import rasterizer as R  # noqa: E402
from tests import gpu_util as G  # noqa: E402
from tests.helpers import small_case  # noqa: E402

for (P, H, W, deg, cs, seed) in [(3000, 100, 75, 4, 9.0, 2), (1400, 32, 32, 1, 60.0, 7), (4000, 32, 32, 0, 60.0, 8),
                                 (517, 48, 48, 3, 4.0, 3)]:
    _, ri = small_case(P, H, W, deg, seed=seed, cov_scale=cs)
    st = G.run_cuda_forward(ri)
    g = torch.tensor(np.random.default_rng(0).standard_normal((3, H, W)).astype(np.float32), device="cuda:0")
    out = R.backward_raw(st, g)
    comp = R.backward_raw(st, g, compact=True)  # the compact (multi-GPU / multi-view) mode of the per-Gaussian kernel
    torch.cuda.synchronize()
    print("case", P, H, W, deg, "N", st["N"], "max", st["max_tile_pairs"], float(out["dmeans3D"].abs().sum()))
