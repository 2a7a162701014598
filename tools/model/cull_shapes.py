"""Surviving (pixel block, Gaussian) fraction of the exact ellipse-vs-rectangle cull for different block shapes on the
C2 scene (CPU, oracle preprocess + binning).  Design aid for DESIGN.md section 8."""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from ggrt_official_b200.synthetic import make_scene, to_raster_inputs
from oracle import c_oracle as co
P,H,W = 300000, 756, 1008
ri = to_raster_inputs(make_scene(P,H,W,sh_degree=4,seed=3407))
cam = co.Camera(W=W,H=H,tanfovx=ri.tanfovx,tanfovy=ri.tanfovy,view=ri.viewmatrix,proj=ri.projmatrix,campos=ri.campos,bg=ri.bg,deg=4)
pre = co.preprocess(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
b = co.bin_tiles(cam, pre)
N=b['N']
gx,gy = cam.grid
pl = b['point_list'].astype(np.int64); ranges=b['ranges'].astype(np.int64)
tile_of = np.zeros(N,np.int64)
for t in range(gx*gy): tile_of[ranges[t,0]:ranges[t,1]] = t
xy = pre['xy'][pl].astype(np.float64); co4 = pre['conic_opacity'][pl].astype(np.float64)
A,B,Cc,o = co4[:,0],co4[:,1],co4[:,2],co4[:,3]
tau = 2*np.log(255*o)
tx = (tile_of % gx)*16; ty=(tile_of//gx)*16
def frac(bw,bh):
    tot=0
    for by in range(16//bh):
        for bx in range(16//bw):
            lox = tx+bx*bw - xy[:,0]; hix = lox+bw-1; loy = ty+by*bh-xy[:,1]; hiy=loy+bh-1
            dxe = np.minimum(np.maximum(0,lox),hix); dye=np.minimum(np.maximum(0,loy),hiy)
            m = (dxe==0)&(dye==0)
            dx = np.clip(-B*dye/A, lox, hix); q1 = A*dx*dx+2*B*dx*dye+Cc*dye*dye
            dy = np.clip(-B*dxe/Cc, loy, hiy); q2 = A*dxe*dxe+2*B*dxe*dy+Cc*dy*dy
            q = np.where(m,0,np.minimum(np.where(dye!=0,q1,3e38),np.where(dxe!=0,q2,3e38)))
            tot += (q<=tau).sum()
    nb = (16//bw)*(16//bh)
    return tot/(N*nb), tot/N*bw*bh
for bw,bh in [(16,16),(16,8),(8,8),(16,4),(8,4),(4,4),(8,2),(4,2),(2,2),(16,2),(16,1),(8,1)]:
    f,pp = frac(bw,bh)
    print(f"block {bw}x{bh}: survive frac {f:.3f}; evaluated px per pair {pp:.1f}; surviving (block,pair) per pair {pp/(bw*bh):.2f}")
