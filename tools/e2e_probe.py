import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from ggrt_official_b200 import GaussianRasterizationSettings, GaussianRasterizer
from ggrt_official_b200 import rasterizer as R
dev = torch.device("cuda:0")
ri, g_np = bench.make_inputs("c2", 0)
t = lambda a: torch.tensor(np.asarray(a), device=dev)
rs = GaussianRasterizationSettings(image_height=ri.image_height, image_width=ri.image_width, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(ri.bg), scale_modifier=1.0, viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix), sh_degree=ri.sh_degree, campos=t(ri.campos), prefiltered=False)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = [pin(ri.means3D), pin(ri.cov3D), pin(ri.opacities), pin(ri.shs)]
grad_img = t(g_np); rast = GaussianRasterizer(rs)
H, W = ri.image_height, ri.image_width
h_img = torch.empty((3, H, W)).pin_memory()
d_in = [torch.empty_like(x, device=dev).requires_grad_() for x in h]
def sync(): torch.cuda.synchronize(); return time.perf_counter()
print("== raw path on the same tensors")
for it in range(4):
    t0 = sync()
    st = R.forward_raw(d_in[0].detach(), d_in[3].detach(), None, d_in[2].detach(), d_in[1].detach(), rs)
    t1 = sync()
    R.backward_raw(st, grad_img)
    t2 = sync()
    print("raw fwd %.3f bwd %.3f" % ((t1-t0)*1e3, (t2-t1)*1e3))
print("== autograd path")
ms = lambda: {k: torch.cuda.memory_stats()[k] for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "reserved_bytes.all.current", "allocated_bytes.all.current")}
for it in range(12):
    print(ms())
    t0 = sync()
    with torch.no_grad():
        for d_, h_ in zip(d_in, h):
            d_.copy_(h_, non_blocking=True); d_.grad = None
    t1 = sync()
    m, c, o, s = d_in
    m2 = torch.zeros_like(m, requires_grad=True)
    ta = sync()
    image, radii, _ = rast(means3D=m, means2D=m2, shs=s, colors_precomp=None, opacities=o, cov3D_precomp=c)
    t2 = sync()
    loss = (image * grad_img).sum()
    t3 = sync()
    loss.backward()
    t4 = sync()
    h_img.copy_(image.detach(), non_blocking=True)
    t5 = sync()
    print("h2d %.2f zeros %.2f fwd %.2f loss %.2f bwd %.2f d2h %.2f ms" % tuple(1e3 * x for x in (t1 - t0, ta - t1, t2 - ta, t3 - t2, t4 - t3, t5 - t4)))
