"""Decoder-level timing at C2: the reference's two-pass flow (colour pass + depth pass, each a full
rasterization) vs the fused single-rasterization path, fwd+bwd."""
import json
import sys
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
out = {}
for name, fused, fast in (("two_pass", False, False), ("fused_depth", True, False), ("fast_glue", True, True)):
    dec = DecoderSplattingCUDA(fused_depth=fused, fast_glue=fast)

    def step():
        for v in leaves.values():
            v.grad = None
        r = dec(Gaussians(**leaves), extr, intr, near, far, (H, W), depth_mode="depth")
        ((r.color * wc).sum() + (r.depth * wd).sum()).backward()
        return r

    for _ in range(5):
        r = step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for a, b in ev:
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    out[name] = dict(ms_median=ms[len(ms) // 2], ms_min=ms[0])
    out[name + "_depth_mean"] = float(r.depth.mean())
out["speedup_fused"] = out["two_pass"]["ms_median"] / out["fused_depth"]["ms_median"]
out["speedup_fast_glue"] = out["two_pass"]["ms_median"] / out["fast_glue"]["ms_median"]
print(json.dumps(out))
