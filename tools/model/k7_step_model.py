"""CPU model of the render-backward kernel's control flow on the C2 scene (oracle binning + the kernel's cull
rule): counts chunks and pixel steps of the current design (matches ncu's executed counts exactly: 186 347 / 1 490 776)
and of alternative sub-block queue layouts.  Design aid for DESIGN.md section 8; uses the oracle, never the product."""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from ggrt_official_b200.synthetic import make_scene, to_raster_inputs
from oracle import c_oracle as co
P,H,W = 300000, 756, 1008
ri = to_raster_inputs(make_scene(P,H,W,sh_degree=4,seed=3407))
cam = co.Camera(W=W,H=H,tanfovx=ri.tanfovx,tanfovy=ri.tanfovy,view=ri.viewmatrix,proj=ri.projmatrix,campos=ri.campos,bg=ri.bg,deg=4)
pre = co.preprocess(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
b = co.bin_tiles(cam, pre); img = co.render_forward(cam, pre, b)
N=b['N']; gx,gy = cam.grid
pl = b['point_list'].astype(np.int64); ranges=b['ranges'].astype(np.int64)
tile_of = np.zeros(N,np.int64)
for t in range(gx*gy): tile_of[ranges[t,0]:ranges[t,1]] = t
pos = np.arange(N)-ranges[tile_of,0]
xy = pre['xy'][pl].astype(np.float64); co4 = pre['conic_opacity'][pl].astype(np.float64)
A,B,Cc,o = co4[:,0],co4[:,1],co4[:,2],co4[:,3]
tau = 2*np.log(255*o)*1.001+0.02
tx = (tile_of % gx)*16; ty=(tile_of//gx)*16
ncon = img['n_contrib'].astype(np.int64)
def hits(x0,y0,w,h):
    lox = x0 - xy[:,0]; hix = lox+w; loy = y0-xy[:,1]; hiy=loy+h
    dxe = np.minimum(np.maximum(0,lox),hix); dye=np.minimum(np.maximum(0,loy),hiy)
    m = (dxe==0)&(dye==0)
    dx = np.clip(-B*dye/A, lox, hix); q1 = A*dx*dx+2*B*dx*dye+Cc*dye*dye
    dy = np.clip(-B*dxe/Cc, loy, hiy); q2 = A*dxe*dxe+2*B*dxe*dy+Cc*dy*dy
    q = np.where(m,0,np.minimum(np.where(dye!=0,q1,3e38),np.where(dxe!=0,q2,3e38)))
    return q<=tau
# warp_last per (tile, block): max n_contrib over the block's pixels
Hp=gy*16; Wp=gx*16
nc = np.zeros((Hp,Wp),np.int64); nc[:H,:W]=ncon
T=gx*gy
steps_now=0; steps_half=0; chunks_now=0; chunks_half=0; cand2=0
for wb in range(8):
    bx=(wb&1)*8; by=(wb>>1)*4
    blk = nc.reshape(gy,16,gx,16)[:,by:by+4,:,bx:bx+8].max(axis=(1,3)).reshape(-1)   # warp_last per tile
    wl = blk[tile_of]
    live = pos < wl
    h = hits(tx+bx, ty+by, 7,3) & live
    qn = np.bincount(tile_of[h], minlength=T)
    c = (qn+7)//8
    chunks_now += c.sum(); steps_now += (c*8).sum()
    cand2 += h.sum()
    for half in range(2):
        hh = h & hits(tx+bx+4*half, ty+by, 3,3)
        qh = np.bincount(tile_of[hh], minlength=T)
        ch = (qh+7)//8
        chunks_half += ch.sum(); steps_half += (ch*4).sum()
print('now: chunks',chunks_now,'steps',steps_now)
print('half queues: chunks',chunks_half,'steps',steps_half, 'second-level candidates',cand2)
s_now = steps_now*64 + chunks_now*85; s_half = steps_half*64 + chunks_half*85 + cand2*3
print('instr model (steps*64 + chunks*85 [+3/candidate]):', s_now/1e6, s_half/1e6, 'ratio', s_half/s_now)
def eval_split(name, subrects, groups_per_sub):
    steps=0; chunks=0
    for wb in range(8):
        bx=(wb&1)*8; by=(wb>>1)*4
        blk = nc.reshape(gy,16,gx,16)[:,by:by+4,:,bx:bx+8].max(axis=(1,3)).reshape(-1)
        live = pos < blk[tile_of]
        h = hits(tx+bx, ty+by, 7,3) & live
        for (ox,oy,w,hgt) in subrects:
            hh = h & hits(tx+bx+ox, ty+by+oy, w,hgt)
            qh = np.bincount(tile_of[hh], minlength=T)
            ch=(qh+7)//8; chunks+=ch.sum(); steps+=(ch*groups_per_sub).sum()
    s = steps*64 + chunks*85 + cand2*3*len(subrects)/2
    print(name,'chunks',chunks,'steps',steps,'model Minstr',s/1e6,'ratio',s/s_now)
eval_split('rows 8x1', [(0,r,7,0) for r in range(4)], 2)
eval_split('quarters 4x2', [(4*a,2*b_,3,1) for a in range(2) for b_ in range(2)], 2)
eval_split('top/bottom 8x2', [(0,0,7,1),(0,2,7,1)], 4)
# per-chunk pixel-group skipping in the CURRENT design: a step (chunk, half-row group) is needed only if one of the
# chunk's 8 Gaussians reaches the group's 4x1 pixels
need=0; tot=0
for wb in range(8):
    bx=(wb&1)*8; by=(wb>>1)*4
    blk = nc.reshape(gy,16,gx,16)[:,by:by+4,:,bx:bx+8].max(axis=(1,3)).reshape(-1)
    live = pos < blk[tile_of]
    h = hits(tx+bx, ty+by, 7,3) & live
    idx = np.nonzero(h)[0]                      # ascending pos within tile; queue order is descending, chunking symmetric enough:
    # emulate exact queue order: per tile descending pos
    order = np.lexsort((-pos[idx], tile_of[idx])); idx = idx[order]
    t_of = tile_of[idx]
    first = np.r_[True, t_of[1:]!=t_of[:-1]]
    start = np.maximum.accumulate(np.where(first, np.arange(len(idx)), 0))
    rank = np.arange(len(idx)) - start
    chunk_id = np.cumsum(first | (rank%8==0)) - 1
    nchunks = chunk_id.max()+1
    for g in range(8):
        gx0 = bx + (g&1)*4; gy0 = by + (g>>1)
        hg = hits(tx+gx0, ty+gy0, 3,0)[idx]
        needed = np.zeros(nchunks,bool); np.logical_or.at(needed, chunk_id, hg)
        need += needed.sum(); tot += nchunks
print('group-skip: needed steps',need,'of',tot, need/tot, 'model ratio', (need*64+ (tot/8)*135)/s_now)
