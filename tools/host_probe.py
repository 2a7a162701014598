"""Host cost of one decoder-level call (device glue, 1 view, colour + depth, fwd+bwd through autograd) at C2: CUDA-event
time per call and the host's enqueue time per call.  Environment toggles (GGRT_RASTER_PDL, GGRT_RASTER_OVERLAP) are read
by the library; run once per setting."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ggrt_official_b200.decoder import DecoderSplattingCUDA, Gaussians  # noqa: E402
from ggrt_official_b200.synthetic import make_scene  # noqa: E402

dev = "cuda:0"
P, H, W = 300_000, 756, 1008
sc = make_scene(P, H, W, sh_degree=4)
t = lambda a: torch.tensor(np.asarray(a), device=dev)
leaves = dict(means=t(sc.means)[None].requires_grad_(), covariances=t(sc.covariances)[None].requires_grad_(),
              harmonics=t(sc.harmonics)[None].requires_grad_(), opacities=t(sc.opacities)[None].requires_grad_())
extr, intr = t(sc.extrinsics)[None, None], t(sc.intrinsics)[None, None]
near, far = torch.full((1, 1), sc.near, device=dev), torch.full((1, 1), sc.far, device=dev)
wc = torch.randn(1, 1, 3, H, W, device=dev) / (3 * H * W)
wd = torch.randn(1, 1, H, W, device=dev) / (H * W)
dec = DecoderSplattingCUDA(device_glue=True)


def step():
    for v in leaves.values():
        v.grad = None
    r = dec(Gaussians(**leaves), extr, intr, near, far, (H, W), depth_mode="depth")
    ((r.color * wc).sum() + (r.depth * wd).sum()).backward()


for _ in range(10):
    step()
torch.cuda.synchronize()
n = 50
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
t0 = time.perf_counter()
for _ in range(n):
    step()
host = (time.perf_counter() - t0) / n
b.record()
torch.cuda.synchronize()
cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")]
print(json.dumps({"env": {k: os.environ.get(k) for k in ("GGRT_RASTER_PDL", "GGRT_RASTER_OVERLAP")},
                  "gpu_ms_per_call": a.elapsed_time(b) / n, "host_ms_per_call": host * 1e3, "cpu": cpu[0] if cpu else None,
                  "ncpu": len(cpu)}))
