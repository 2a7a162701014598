"""Forward / backward wall time (CUDA events) vs the sum of their kernel times: how big are the gaps?"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from ggrt_official_b200 import GaussianRasterizationSettings, _cabi
from ggrt_official_b200 import rasterizer as R
dev = torch.device("cuda:0")
ri, g_np = bench.make_inputs("c2", 0)
t = lambda a: torch.tensor(np.asarray(a), device=dev)
rs = GaussianRasterizationSettings(image_height=ri.image_height, image_width=ri.image_width, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(ri.bg), scale_modifier=1.0, viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix), sh_degree=ri.sh_degree, campos=t(ri.campos), prefiltered=False)
args = (t(ri.means3D), t(ri.shs), None, t(ri.opacities), t(ri.cov3D), rs)
g = t(g_np)
for _ in range(5):
    st = R.forward_raw(*args); R.backward_raw(st, g)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
fw, bw = [], []
for _ in range(30):
    ev[0].record(); st = R.forward_raw(*args); ev[1].record(); R.backward_raw(st, g); ev[2].record()
    torch.cuda.synchronize()
    fw.append(ev[0].elapsed_time(ev[1])); bw.append(ev[1].elapsed_time(ev[2]))
_cabi.profile_enable(True)
kf, kb = [], []
for _ in range(10):
    st = R.forward_raw(*args); a = _cabi.profile_read(); R.backward_raw(st, g); b = _cabi.profile_read()
    kf.append(sum(a[k] for k in ("geometry", "scan_tiles", "color", "emit", "sort_tiles", "render_forward")))
    kb.append(b["render_backward"] + b["preprocess_backward"])
print("forward wall %.1f us, kernels %.1f us | backward wall %.1f us, kernels %.1f us" % (1e3*np.median(fw), 1e3*np.median(kf), 1e3*np.median(bw), 1e3*np.median(kb)))
