import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ggrt_official_b200.decoder import DecoderSplattingCUDA, Gaussians
from ggrt_official_b200.synthetic import make_scene
from torch.profiler import profile, ProfilerActivity
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
dec = DecoderSplattingCUDA(fused_depth=True)
def step():
    for v in leaves.values(): v.grad = None
    r = dec(Gaussians(**leaves), extr, intr, near, far, (H, W), depth_mode="depth")
    ((r.color * wc).sum() + (r.depth * wd).sum()).backward()
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
