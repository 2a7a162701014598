// Per-tile back-to-front gradient replay (A.4).  Replaces upstream renderCUDA (bwd),
// SURVEY.md 8a row a12 -- the dominant cost of fwd+bwd upstream because every pixel
// issues ~10 global float atomics per contributing Gaussian.
//
// Work decomposition ("Gaussian-parallel"): one CTA per 16x16 tile, warp w owns the 8x4
// pixel block (w&1, w>>1) as in the forward.  A batch of Gaussian records is staged into shared
// memory with cp.async together with the forward kernel's cull result (one bit per warp pixel
// block: exact ellipse vs rectangle test on {alpha >= 1/255}); each warp compacts the entries
// that reach its block into a back-to-front queue and then consumes the queue 16 Gaussians
// at a time.  LANES = (Gaussian pair k = lane >> 2, pixel slot q = lane & 3): per step (one row of
// the block) a lane evaluates ITS two Gaussians (queue entries 2k, 2k+1) at ITS two pixels
// (q, row) and (q + 4, row), so the warp covers 16 Gaussians x 8 pixels.
//
//  * Alpha-blend replay as a prefix scan ("warp-shuffle prefix scans for the alpha-blend", north_star).  Going back
//    to front each Gaussian maps the running pair (T, -R) to (a T, -R + nb T), a = 1/(1-alpha),
//    nb = -alpha (c.g)/(1-alpha).  These maps compose associatively: a lane composes its two Gaussians in
//    registers, ONE 3-level warp scan over the 8 pairs (shuffle distance 4, 8, 16) yields every pair's inclusive
//    prefix, and the prefix of the pair's back Gaussian follows by undoing the front one (multiply by 1 - alpha).
//    Lanes without a partner at some level blend in the identity map through per-lane 0/1 masks (no selects).
//        dL/dalpha_i = [ T_i (c_i . g) - sum_{j behind i, j included} w_j (c_j . g) - T_final (bg . g) ] / (1 - alpha_i)
//  * Gradient accumulation on the tensor pipe.  The nine sums a Gaussian needs over the pixels of the block,
//        sum_p q_ip {1, x_p, y_p, x_p^2, x_p y_p, y_p^2}   (q = G dL/dalpha: opacity, mean and conic gradients)
//        sum_p w_ip {g_r, g_g, g_b, g_aux}_p               (w = alpha T: colour gradients)
//    are [16 Gaussians x 8 pixels] . [8 pixels x 8 features] products: mma.sync.m16n8k8 (tf32 inputs, fp32
//    accumulate; SASS HMMA.1688.F32.TF32), whose A fragment is exactly this lane layout.  Pixel coordinates are taken
//    relative to the block centre (+-3.5, +-1.5 and their products: exact in tf32); q, w and the pixel gradients are
//    split into a tf32 head and a remainder (x = hi + lo, two MMAs), so the sums carry ~22 mantissa bits.  The D
//    fragments replace the per-lane register accumulators AND the per-chunk shuffle fold of the first version; they
//    are converted to gradients w.r.t. the Gaussian's own centre once per chunk (binomial shift by u = gx - xc) by
//    16 lanes and committed with vector RED.ADD.F32 (2 x v4 + 1 scalar per Gaussian per warp).
// Per-pixel running state {T behind, -(sum behind)} lives in shared memory between chunks.
//
// The kernel is instruction-issue bound (ncu: DRAM 3 %).  Everything per-pixel-pair runs on packed FP32 pairs
// (FFMA2 / FMUL2 / FADD2, f32x2.cuh): the two pixels of a lane share dy and their dx differ by the constant 4.
// Signs are arranged so that no negation is ever issued: the opacity is negated once per chunk (n_alpha = -alpha
// falls out of the multiply), the state carries -R, and the colour sums come out negated and are fixed at commit.
// History and measurements: DESIGN.md section 4 (K7).
#include "f32x2.cuh"
#include "render_common.cuh"

namespace ggrt {

#ifndef GGRT_BWD_MMA
#define GGRT_BWD_MMA 1
#endif
#ifndef GGRT_BWD_ROWPAIR
#define GGRT_BWD_ROWPAIR 0
#endif
#ifndef GGRT_BWD_BATCH
#define GGRT_BWD_BATCH (GGRT_BWD_MMA ? 384 : 512)
#endif
#ifndef GGRT_BWD_MINBLOCKS
#define GGRT_BWD_MINBLOCKS (GGRT_BWD_MMA ? 3 : 4)
#endif
constexpr int BWD_BATCH = GGRT_BWD_BATCH;
constexpr int NWARPS = BWD_WARPS;
#ifndef GGRT_BWD_ROW_UNROLL
#define GGRT_BWD_ROW_UNROLL 2
#endif
constexpr int ROW_UNROLL = GGRT_BWD_ROW_UNROLL;
constexpr int GL = 8;  // (first version) Gaussians per warp step
constexpr int QL = 4;  // pixel slots per warp step (each slot = the pixel pair (q, q + 4) of one row)

__device__ __forceinline__ void red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// 16-byte vector reduction (sm_90+): one L2 operation for four adjacent floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ f2 lds_f2(uint32_t a) {
    f2 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v.v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void lds_2f2(uint32_t a, f2& x, f2& y) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x.v), "=l"(y.v) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_2f2(uint32_t a, f2 x, f2 y) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a), "l"(x.v), "l"(y.v) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t a, f2 x) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(x.v) : "memory");
}

// One level of the segmented scan over the GL = 8 Gaussian slots (consecutive lanes): (A, nB) <- (A Ap, nB Ap + nBp)
// where (Ap, nBp) come from the lane d slots further back; lanes without such a lane keep their values (the shuffle's
// in-range predicate drives the two packed operations, no compare, no select).  [first version]
__device__ __forceinline__ void scan_step(f2& A, f2& nB, int d, int gl) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 alo, ahi, blo, bhi;\n\t"
        ".reg .b64 ap, bp;\n\t"
        "mov.b64 {alo, ahi}, %0;\n\t"
        "mov.b64 {blo, bhi}, %1;\n\t"
        "setp.ge.s32 p, %3, %2;\n\t"
        "shfl.sync.up.b32 alo, alo, %2, 0x1800, 0xffffffff;\n\t"
        "shfl.sync.up.b32 ahi, ahi, %2, 0x1800, 0xffffffff;\n\t"
        "shfl.sync.up.b32 blo, blo, %2, 0x1800, 0xffffffff;\n\t"
        "shfl.sync.up.b32 bhi, bhi, %2, 0x1800, 0xffffffff;\n\t"
        "mov.b64 ap, {alo, ahi};\n\t"
        "mov.b64 bp, {blo, bhi};\n\t"
        "@p fma.rn.f32x2 %1, %1, ap, bp;\n\t"
        "@p mul.rn.f32x2 %0, %0, ap;\n\t"
        "}"
        : "+l"(A.v), "+l"(nB.v)
        : "r"(d), "r"(gl));
}

}  // namespace ggrt

#if !GGRT_BWD_MMA
#include "render_bwd_v1.cuh"
#else

namespace ggrt {

constexpr int GC = 16;            // Gaussians per warp step / chunk
constexpr int STAGE_STRIDE = 20;  // floats per Gaussian row of the per-warp D staging area (16 used, 80-byte rows)
constexpr uint32_t TF32_MASK = 0xffffe000u;

// D[16x8] += A[16x8] . B[8x8], tf32 inputs (the low 13 mantissa bits of the registers are ignored), fp32 accumulate.
// A: a0 = (row g, col c), a1 = (g + 8, c), a2 = (g, c + 4), a3 = (g + 8, c + 4); B: b0 = (c, g), b1 = (c + 4, g);
// D: d0, d1 = (g, 2c), (g, 2c + 1); d2, d3 = (g + 8, 2c), (g + 8, 2c + 1); with g = lane >> 2, c = lane & 3.
__device__ __forceinline__ void mma_tf32(float (&d)[4], float a0, float a1, float a2, float a3, float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a0)), "r"(__float_as_uint(a1)), "r"(__float_as_uint(a2)), "r"(__float_as_uint(a3)),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ float tf32_head(float x) { return __uint_as_float(__float_as_uint(x) & TF32_MASK); }
__device__ __forceinline__ f2 tf32_head2(f2 x) { return pk(tf32_head(lo(x)), tf32_head(hi(x))); }
// remainder of a packed pair after its tf32 heads h: x - h, exact
__device__ __forceinline__ f2 tf32_tail2(f2 x, f2 h) { return fma2(h, bc(-1.0f), x); }

// One level of the scan over the 8 Gaussian pairs (lane distance 4 d): (A, nB) <- (A Ap, nB Ap + nBp) with (Ap, nBp)
// from the lane d pairs further back; m = 1 where such a lane exists, else 0 (om = 1 - m): the identity map.
__device__ __forceinline__ void pair_scan_step(f2& A, f2& nB, int delta, f2 m, f2 om) {
    f2 ap, bp;
    ap.v = __shfl_up_sync(0xffffffffu, A.v, delta);
    bp.v = __shfl_up_sync(0xffffffffu, nB.v, delta);
    ap = fma2(ap, m, om);
    bp = mul2(bp, m);
    nB = fma2(nB, ap, bp);
    A = mul2(A, ap);
}

struct GaussReg {  // per-chunk constants of one of a lane's two Gaussians
    f2 dx2, pa2;     // dx at the lane's two pixels; ea dx^2
    float eb, ec, gyr, ncw;
    float cr, cg, cb, cx;
    uint32_t pos;
};

struct Eval {  // one Gaussian at the lane's pixel pair of one row
    f2 G2, nal2, om2, io2, sdot2;
    bool acta, actb;
};

template <bool AUX>
__device__ __forceinline__ void load_gauss(GaussReg& r, uint32_t sbase, uint32_t jj, bool valid, uint32_t boff, float xq,
                                           float by0f) {
    const uint32_t src = sbase + jj * REC_BYTES;
    const float2 gxy = lds64(src);
    float4 con = lds128(src + 16);
    const float4 col = lds128(src + 32);
    if (!valid) con.w = 0.f;  // zero opacity: never active
    r.pos = valid ? boff + jj : 0xffffffffu;  // 0-based list position (never "behind" a pixel's last)
    asm volatile("" : "+r"(r.pos));           // keep the select: otherwise `valid` is re-tested in every pixel step
    // exponent of the Gaussian in base 2 with the -1/2 folded in: G = 2^(ea dx^2 + eb dx dy + ec dy^2)
    const float ea = -0.5f * LOG2E * con.x;
    r.eb = -LOG2E * con.y, r.ec = -0.5f * LOG2E * con.z;
    r.ncw = -con.w;
    const float dxa = gxy.x - xq;
    r.dx2 = pk(dxa, dxa - 4.0f);
    r.pa2 = mul2(bc(ea), mul2(r.dx2, r.dx2));
    r.gyr = gxy.y - by0f;
    r.cr = col.x, r.cg = col.y, r.cb = col.z, r.cx = AUX ? col.w : 0.f;
}

template <bool AUX>
__device__ __forceinline__ void eval_gauss(Eval& e, const GaussReg& r, float rowf, f2 gr2, f2 gg2, f2 gb2, f2 ga2,
                                           f2 last2) {
    const float dy = r.gyr - rowf;
    // log2 of the Gaussian at both pixels: ea dx^2 + (eb dx + ec dy) dy
    const f2 pw2 = fma2(fma2(bc(r.eb), r.dx2, bc(r.ec * dy)), bc(dy), r.pa2);
    const float pwa = lo(pw2), pwb = hi(pw2);
    e.G2 = pk(ex2_approx(pwa), ex2_approx(pwb));
    const f2 arn2 = mul2(bc(r.ncw), e.G2);  // -(opacity * G)
    const float arna = fmaxf(-ALPHA_MAX, lo(arn2)), arnb = fmaxf(-ALPHA_MAX, hi(arn2));
    e.acta = (r.pos < __float_as_uint(lo(last2))) && (pwa <= 0.0f) && (arna <= -ALPHA_MIN);
    e.actb = (r.pos < __float_as_uint(hi(last2))) && (pwb <= 0.0f) && (arnb <= -ALPHA_MIN);
    e.nal2 = pk(e.acta ? arna : 0.f, e.actb ? arnb : 0.f);  // -alpha
    e.om2 = add2(bc(1.0f), e.nal2);                          // 1 - alpha
    e.io2 = pk(rcp_approx(lo(e.om2)), rcp_approx(hi(e.om2)));
    e.sdot2 = fma2(bc(r.cb), gb2, fma2(bc(r.cg), gg2, mul2(bc(r.cr), gr2)));
    if (AUX) e.sdot2 = fma2(bc(r.cx), ga2, e.sdot2);
}

// Per-warp shared memory (one block per warp, so that two base registers + immediate offsets reach everything):
//   pix[j], j = row * 4 + q (pixels (q, row) = "A" and (q + 4, row) = "B"):
//          {g_r A, g_r B, g_g A, g_g B | g_b A, g_b B, last A (bits), last B (bits) |
//           T behind A, T behind B, -(sum behind) A, -(sum behind) B}
//   ga[j] = {g_aux A, g_aux B}
//   fcm[row][lane] = {c0, c1, m0, m1}: the lane's two B fragments for that row.  {c0, c1}: pixel-gradient feature
//          n = lane >> 2 at pixels q = lane & 3 and q + 4 (n = 0..3: tf32 heads of g_r, g_g, g_b, g_aux, n = 4..7: their
//          remainders); {m0, m1}: coordinate monomial n of {1, x, y, x^2, xy, y^2, 0, 0} at the same pixels,
//          x = column - 3.5, y = row - 1.5
//   stage[16][STAGE_STRIDE]: the chunk's D fragments as per-Gaussian rows;  queue: the culled batch, back to front
struct WarpShared {
    float pix[16][12];
    float ga[16][2];
    float fcm[4][32][4];
    float stage[GC][STAGE_STRIDE];
    unsigned short queue[BWD_BATCH];
};
static_assert(sizeof(WarpShared) % 16 == 0, "per-warp blocks stay 16-byte aligned");

// AUX: the 4th blended channel (out_depth) also carries an upstream gradient.
template <bool AUX>
__global__ void __launch_bounds__(BWD_THREADS, GGRT_BWD_MINBLOCKS * 8 / BWD_WARPS)
render_backward_kernel(View v, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                       const float4* __restrict__ rec2, const uint32_t* __restrict__ starts,
                       const uint32_t* __restrict__ points, const uint8_t* __restrict__ masks,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dout, const float* __restrict__ dL_dout_aux,
                       float* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];  // WarpShared[NWARPS] | records of the batch
    __shared__ uint32_t sid[BWD_BATCH];
    static_assert(BWD_BATCH % 4 == 0, "the compaction reads the masks four at a time");
    __shared__ __align__(4) uint8_t smask[BWD_BATCH];  // the forward kernel's cull result per staged record (bit = warp pixel block)
    __shared__ uint32_t block_last_s;

    WarpShared* const wsm = reinterpret_cast<WarpShared*>(dyn_smem);
    const uint32_t sbase = smem_addr(dyn_smem + NWARPS * sizeof(WarpShared));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WarpShared& ws = wsm[warp];
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int wt = warp + blockIdx.z * BWD_WARPS;  // warp pixel block of the tile (8 per tile)
    const int bx0 = blockIdx.x * TILE + (wt & 1) * 8, by0 = blockIdx.y * TILE + (wt >> 1) * 4;
    const float bx0f = (float)bx0, by0f = (float)by0;
    // The per-pixel loads are issued first: they do not depend on the tile's list range, so their latency overlaps
    // the dependent chain  tile range -> list entries -> records  of the speculative gather below.
    const int lx = lane & 7, ly = lane >> 3;
    const int px = bx0 + lx, py = by0 + ly;
    uint32_t last = 0;
    float Tfin = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, da = 0.f;
    pdl_enter();  // (programmatic dependent launch: the CTA is resident before the forward render kernel has drained)
    if (px < v.W && py < v.H) {
        const size_t pix = (size_t)py * v.W + px, hw = (size_t)v.H * v.W;
        Tfin = final_T[pix];
        last = n_contrib[pix];
        d0 = dL_dout[pix];
        d1 = dL_dout[hw + pix];
        d2 = dL_dout[2 * hw + pix];
        if (AUX) da = dL_dout_aux[pix];
    }
    const uint32_t start = starts[tile], tile_cnt = starts[tile + 1] - start;

    // Records of a batch of list entries [boff, boff + cnt) -> shared memory with cp.async (SASS LDGSTS), one group
    auto gather = [&](uint32_t boff, uint32_t cnt) {
        for (uint32_t k = tid; k < cnt; k += BWD_THREADS) {
            const uint32_t id = points[start + boff + k];
            sid[k] = id;
            smask[k] = masks[start + boff + k];
            const uint32_t dst = sbase + k * REC_BYTES;
            cp_async16(dst, rec0 + id);
            cp_async16(dst + 16, rec1 + id);
            cp_async16(dst + 32, rec2 + id);
        }
        cp_async_commit();
    };
    // The batches are consumed back to front and the first one is normally the tile's last (unless every pixel
    // saturated early): start its gather now, so that it overlaps the per-pixel loads and the reduction below.
    const uint32_t spec_b = tile_cnt ? (tile_cnt - 1) / BWD_BATCH : 0;
    gather(spec_b * BWD_BATCH, tile_cnt - spec_b * BWD_BATCH);

    // ---- per-pixel constants / initial state (lane = pixel here) -------------------------------
    {
        // pixels with a zero upstream gradient contribute nothing (crop training leaves most tiles empty)
        if (d0 == 0.f && d1 == 0.f && d2 == 0.f && da == 0.f) last = 0;
        const float bg_dot = v.bg[0] * d0 + v.bg[1] * d1 + v.bg[2] * d2;
        const int qq = lx & 3, h = lx >> 2;  // pixel slot, half (0 = A, 1 = B)
        const int j = ly * 4 + qq;
        float* sp = ws.pix[j];
        sp[h] = d0, sp[2 + h] = d1;
        sp[4 + h] = d2, sp[6 + h] = __uint_as_float(last);
        sp[8 + h] = Tfin, sp[10 + h] = -(Tfin * bg_dot);
        ws.ga[j][h] = da;
        // B fragments of the pixel-gradient features: lane n * 4 + qq of row ly holds feature n at pixels qq, qq + 4
        const float f[4] = {d0, d1, d2, da};
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float head = tf32_head(f[n]);
            ws.fcm[ly][n * 4 + qq][h] = head;
            ws.fcm[ly][(n + 4) * 4 + qq][h] = f[n] - head;
        }
        {  // coordinate monomials relative to the block centre (this lane's fragment): x^ex y^ey, n -> (ex, ey)
            const int n = lane >> 2;
            const int ex = (n == 1 || n == 4) ? 1 : (n == 3 ? 2 : 0), ey = (n == 2 || n == 4) ? 1 : (n == 5 ? 2 : 0);
            const float keep = n < 6 ? 1.f : 0.f;
            const float xa = (float)(lane & 3) - 3.5f, xb = xa + 4.0f;
            const float fxa = keep * (ex == 0 ? 1.f : ex == 1 ? xa : xa * xa), fxb = keep * (ex == 0 ? 1.f : ex == 1 ? xb : xb * xb);
#pragma unroll
            for (int row = 0; row < 4; ++row) {
                const float y = (float)row - 1.5f;
                const float fy = ey == 0 ? 1.f : ey == 1 ? y : y * y;
                ws.fcm[row][lane][2] = fxa * fy;
                ws.fcm[row][lane][3] = fxb * fy;
            }
        }
    }
    if (tid == 0) block_last_s = 0;
    __syncthreads();
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0 && warp_last > 0) atomicMax(&block_last_s, warp_last);
    __syncthreads();
    const uint32_t block_last = block_last_s;
    if (block_last == 0) {
        cp_async_wait<0>();  // nothing may still be in flight into this CTA's shared memory when it exits
        return;
    }
    const uint32_t pmask = __ballot_sync(0xffffffffu, last > 0);  // pixels of this warp that matter (bit = ly*8 + lx)

    const float neg_half_w = -0.5f * (float)v.W, neg_half_h = -0.5f * (float)v.H;
    const int q = lane & 3, kp = lane >> 2;
    // Base addresses, opaque to the compiler (it would otherwise rebuild them from the thread index in every row):
    //   pq: pix[q] (+ row * 192; ga[q] follows at a fixed distance)    pl: fcm[0][lane] (+ row * 512)
    //   front: non-zero in the lanes of the frontmost pair, which own the state stores (kept as a register: the
    //   compiler would otherwise re-derive the predicate from the thread index in every row)
    uint32_t pq = smem_addr(&ws.pix[q][0]), pl = smem_addr(&ws.fcm[0][lane][0]);
    uint32_t front = kp == 7 ? 1u : 0u;
    uint32_t stage_w = smem_addr(&ws.stage[2 * kp][2 * q]);  // this lane's D elements: rows 2 kp, 2 kp + 1
    asm volatile("" : "+r"(pq), "+r"(pl), "+r"(front), "+r"(stage_w));
    const uint32_t pga = pq + (uint32_t)(offsetof(WarpShared, ga) - offsetof(WarpShared, pix)) - 40u * (uint32_t)q;  // ga[q]
    const float xq = bx0f + (float)q;  // x of this lane's pixel A; pixel B is 4 to the right
    // identity masks of the pair scan: level d needs a lane 4 d below
    const f2 m1 = bc(kp >= 1 ? 1.f : 0.f), m2 = bc(kp >= 2 ? 1.f : 0.f), m4 = bc(kp >= 4 ? 1.f : 0.f);
    const f2 om1 = bc(kp >= 1 ? 0.f : 1.f), om2 = bc(kp >= 2 ? 0.f : 1.f), om4 = bc(kp >= 4 ? 0.f : 1.f);

    const int nb = (int)((block_last + BWD_BATCH - 1) / BWD_BATCH);
    for (int bi = nb - 1; bi >= 0; --bi) {
        const uint32_t boff = (uint32_t)bi * BWD_BATCH;
        const uint32_t cnt = min((uint32_t)BWD_BATCH, block_last - boff);
        if (bi != nb - 1 || (uint32_t)bi != spec_b) {  // (the speculative gather above covers the first batch)
            __syncthreads();  // every warp is done with the previous batch before the refill
            if (bi == nb - 1) cp_async_wait<0>();  // the unused speculative batch has landed before it is overwritten
            gather(boff, cnt);
        }
        cp_async_wait<0>();
        __syncthreads();
        if (warp_last <= boff) continue;

        // ---- the entries that reach this warp's pixel block (the forward kernel's exact ellipse-vs-rectangle cull,
        // one bit per block), compacted into a back-to-front queue ---------------------------------------------
        uint32_t qn = 0;
        const uint32_t lim = min(cnt, warp_last - boff);  // entries at or beyond warp_last never contribute here
        for (int R = (int)((lim - 1) & ~127u); R >= 0; R -= 128) {  // 128 entries per round: four mask bytes per lane
            const uint32_t e0 = (uint32_t)R + 4u * (31u - (uint32_t)lane);  // lane 0 holds the round's backmost four
            uint32_t w = 0;
            if (e0 < lim) {
                w = (*reinterpret_cast<const uint32_t*>(&smask[e0]) >> wt) & 0x01010101u;
                const uint32_t nv = lim - e0;
                if (nv < 4u) w &= (1u << (8u * nv)) - 1u;
            }
            const uint32_t b3 = __ballot_sync(0xffffffffu, w & 0x01000000u), b2 = __ballot_sync(0xffffffffu, w & 0x00010000u),
                           b1 = __ballot_sync(0xffffffffu, w & 0x00000100u), b0 = __ballot_sync(0xffffffffu, w & 0x00000001u);
            const uint32_t lt = (1u << lane) - 1u;
            uint32_t pos = qn + __popc(b3 & lt) + __popc(b2 & lt) + __popc(b1 & lt) + __popc(b0 & lt);
            if (w & 0x01000000u) ws.queue[pos++] = (unsigned short)(e0 + 3u);
            if (w & 0x00010000u) ws.queue[pos++] = (unsigned short)(e0 + 2u);
            if (w & 0x00000100u) ws.queue[pos++] = (unsigned short)(e0 + 1u);
            if (w & 0x00000001u) ws.queue[pos++] = (unsigned short)e0;
            qn += __popc(b3) + __popc(b2) + __popc(b1) + __popc(b0);
        }
        if (lane == 0 && (qn & 1u)) ws.queue[qn] = 0;  // the pair read below never sees stale bits
        __syncwarp();

        // ---- 16 queued Gaussians at a time; lane = (Gaussian pair kp, pixel slot q) ------------------
        const uint32_t qaddr = smem_addr(&ws.queue[2 * kp]);
        for (uint32_t c0 = 0; c0 < qn; c0 += GC) {
            uint32_t jpair = 0;
            const bool valid_e = c0 + 2 * kp < qn, valid_o = c0 + 2 * kp + 1 < qn;
            if (valid_e) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(jpair) : "r"(qaddr + c0 * 2) : "memory");
            GaussReg ge, go;  // e: the pair's back Gaussian (queue entry 2 kp), o: the one in front of it
            load_gauss<AUX>(ge, sbase, jpair & 0xffffu, valid_e, boff, xq, by0f);
            load_gauss<AUX>(go, sbase, jpair >> 16, valid_o, boff, xq, by0f);
            float Dm[4] = {0.f, 0.f, 0.f, 0.f}, Dc[4] = {0.f, 0.f, 0.f, 0.f};

            // One row step in three parts, so that GGRT_BWD_ROWPAIR can interleave the dependent shuffle chains of two rows
            struct RowState {
                Eval e, o;
                f2 A2, nB2, nBo, Tb2, nRb2;
            };
            auto row_front = [&](int row, RowState& r) {  // loads, both Gaussians' evaluations, the pair as one map
                const uint32_t off = pq + (uint32_t)row * 192u;  // an immediate offset once the loop is unrolled
                f2 gr2, gg2, gb2, last2, ga2 = bc(0.f);
                lds_2f2(off, gr2, gg2);
                lds_2f2(off + 16, gb2, last2);
                lds_2f2(off + 32, r.Tb2, r.nRb2);
                if (AUX) ga2 = lds_f2(pga + (uint32_t)row * 32u);
                eval_gauss<AUX>(r.e, ge, (float)row, gr2, gg2, gb2, ga2, last2);
                eval_gauss<AUX>(r.o, go, (float)row, gr2, gg2, gb2, ga2, last2);
                const f2 nBe = mul2(mul2(r.e.nal2, r.e.sdot2), r.e.io2);
                r.nBo = mul2(mul2(r.o.nal2, r.o.sdot2), r.o.io2);
                r.A2 = mul2(r.e.io2, r.o.io2), r.nB2 = fma2(r.nBo, r.e.io2, nBe);  // (a_o a_e, nb_o a_e + nb_e)
            };
            auto row_back = [&](int row, RowState& r) {
                const Eval &e = r.e, &o = r.o;
                const f2 A2 = r.A2, nB2 = r.nB2, Tb2 = r.Tb2, nRb2 = r.nRb2;
                // prefix through o = the scan's result (A, nB) = (a_o A', nb_o A' + nB'); prefix through e = (A', nB'):
                // undo o with 1 / a_o = 1 - alpha_o
                const f2 Ae = mul2(A2, o.om2), nBi = fma2(mul2(r.nBo, bc(-1.0f)), Ae, nB2);
                const f2 Ti_o = mul2(Tb2, A2), nRt_o = fma2(Tb2, nB2, nRb2);  // transmittance in front of o; -(sum behind, o incl.)
                const f2 Ti_e = mul2(Tb2, Ae), nRt_e = fma2(Tb2, nBi, nRb2);
                if (front) sts_2f2(pq + 32u + (uint32_t)row * 192u, Ti_o, nRt_o);  // frontmost pair: the state behind the next chunk
                const f2 dq_e = mul2(fma2(Ti_e, e.sdot2, nRt_e), e.io2), dq_o = mul2(fma2(Ti_o, o.sdot2, nRt_o), o.io2);  // dL/dalpha
                // A fragments {e at A, o at A, e at B, o at B}: built with scalar operations, so that each lands in the
                // register the MMA wants (pairing the packed results would cost a move per element)
                const float qAe = e.acta ? lo(e.G2) * lo(dq_e) : 0.f, qAo = o.acta ? lo(o.G2) * lo(dq_o) : 0.f;  // G dL/dalpha
                const float qBe = e.actb ? hi(e.G2) * hi(dq_e) : 0.f, qBo = o.actb ? hi(o.G2) * hi(dq_o) : 0.f;
                const float wAe = lo(e.nal2) * lo(Ti_e), wAo = lo(o.nal2) * lo(Ti_o);  // -(blend weight)
                const float wBe = hi(e.nal2) * hi(Ti_e), wBo = hi(o.nal2) * hi(Ti_o);
                // tf32 heads (masked explicitly: the result does not depend on how the tensor pipe treats the low bits)
                const f2 qA = pk(qAe, qAo), qB = pk(qBe, qBo), wA = pk(wAe, wAo), wB = pk(wBe, wBo);
                const f2 qhA = tf32_head2(qA), qhB = tf32_head2(qB), whA = tf32_head2(wA), whB = tf32_head2(wB);
                const f2 qtA = tf32_tail2(qA, qhA), qtB = tf32_tail2(qB, qhB), wtA = tf32_tail2(wA, whA), wtB = tf32_tail2(wB, whB);
                const float4 fb = lds128(pl + (uint32_t)row * 512u);  // {colour b0, b1, monomial b0, b1}
                mma_tf32(Dm, lo(qhA), hi(qhA), lo(qhB), hi(qhB), fb.z, fb.w);
                mma_tf32(Dm, lo(qtA), hi(qtA), lo(qtB), hi(qtB), fb.z, fb.w);
                mma_tf32(Dc, lo(whA), hi(whA), lo(whB), hi(whB), fb.x, fb.y);
                mma_tf32(Dc, lo(wtA), hi(wtA), lo(wtB), hi(wtB), fb.x, fb.y);
            };
#if GGRT_BWD_ROWPAIR
#pragma unroll
            for (int rp = 0; rp < 4; rp += 2) {
                if (((pmask >> (rp * 8)) & 0xffffu) == 0) continue;  // no pixel of the two rows matters
                RowState r0, r1;
                row_front(rp, r0);
                row_front(rp + 1, r1);
                pair_scan_step(r0.A2, r0.nB2, 4, m1, om1);
                pair_scan_step(r1.A2, r1.nB2, 4, m1, om1);
                pair_scan_step(r0.A2, r0.nB2, 8, m2, om2);
                pair_scan_step(r1.A2, r1.nB2, 8, m2, om2);
                pair_scan_step(r0.A2, r0.nB2, 16, m4, om4);
                pair_scan_step(r1.A2, r1.nB2, 16, m4, om4);
                row_back(rp, r0);
                row_back(rp + 1, r1);
            }
#else
#pragma unroll
            for (int row = 0; row < 4; ++row) {
                if (((pmask >> (row * 8)) & 0xffu) == 0) continue;  // no pixel of the row matters
                RowState r;
                row_front(row, r);
                // the inclusive scan over the pairs
                pair_scan_step(r.A2, r.nB2, 4, m1, om1);
                pair_scan_step(r.A2, r.nB2, 8, m2, om2);
                pair_scan_step(r.A2, r.nB2, 16, m4, om4);
                row_back(row, r);
            }
#endif
            // ---- D fragments -> per-Gaussian rows {S1, Sx, Sy, Sxx, Sxy, Syy, -, - | head sums r g b x | tail sums} ----
            sts_f2(stage_w, pk(Dm[0], Dm[1]));
            sts_f2(stage_w + STAGE_STRIDE * 4, pk(Dm[2], Dm[3]));
            sts_f2(stage_w + 32, pk(Dc[0], Dc[1]));
            sts_f2(stage_w + STAGE_STRIDE * 4 + 32, pk(Dc[2], Dc[3]));
            __syncwarp();
            if (lane < GC && c0 + lane < qn) {
                const uint32_t jj = ws.queue[c0 + lane];
                const uint32_t src = sbase + jj * REC_BYTES;
                const float2 gxy = lds64(src);
                const float4 con = lds128(src + 16);
                const uint32_t stage_r = smem_addr(&ws.stage[lane][0]);
                const float4 s0 = lds128(stage_r), s1 = lds128(stage_r + 16), ch = lds128(stage_r + 32),
                             ct = lds128(stage_r + 48);
                // moments about the block centre -> about the Gaussian's centre: d = u - x with u = g - centre
                const float u = gxy.x - (bx0f + 3.5f), w = gxy.y - (by0f + 1.5f);
                const float S1 = s0.x, Sx = s0.y, Sy = s0.z, Sxx = s0.w, Sxy = s1.x, Syy = s1.y;
                const float mx = fmaf(u, S1, -Sx), my = fmaf(w, S1, -Sy);                 // sum q dx, sum q dy
                const float mA = fmaf(u, mx, fmaf(-u, Sx, Sxx));                          // sum q dx^2
                const float mB = fmaf(w, mx, fmaf(-u, Sy, Sxy));                          // sum q dx dy
                const float mC = fmaf(w, my, fmaf(-w, Sy, Syy));                          // sum q dy^2
                const float a_mx = con.w * mx, a_my = con.w * my;
                float* dst = scratch + (size_t)sid[jj] * GRAD_STRIDE;
                // scratch rows are 48 B (16-B aligned): slots {mx,my,A,B} {C,op,r,g} {b}
                const float gmx = fmaf(con.x, a_mx, con.y * a_my), gmy = fmaf(con.z, a_my, con.y * a_mx);
                const float hw = -0.5f * con.w;
                red_add_v4(dst + G_MX, gmx * neg_half_w, gmy * neg_half_h, hw * mA, hw * mB);
                red_add_v4(dst + G_CC, hw * mC, S1, -(ch.x + ct.x), -(ch.y + ct.y));
                red_add(dst + G_B, -(ch.z + ct.z));
                if (AUX) red_add(dst + G_AUX, -(ch.w + ct.w));
            }
            __syncwarp();
        }
    }
}

}  // namespace ggrt
#endif  // GGRT_BWD_MMA

namespace ggrt {

void launch_render_backward(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, const float* dL_dout,
                            const float* dL_dout_aux, float* scratch, cudaStream_t s) {
    dim3 grid(v.gx, v.gy, 8 / BWD_WARPS);
#if GGRT_BWD_MMA
    constexpr size_t dyn = NWARPS * sizeof(WarpShared) + (size_t)BWD_BATCH * REC_BYTES;
    static const bool opted_in = [] {  // static + dynamic shared memory exceeds the 48 KB default (per device, cheap)
        cudaFuncSetAttribute(render_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        cudaFuncSetAttribute(render_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        return true;
    }();
    (void)opted_in;
#define BWD_MASKS b.masks,
#else
    constexpr size_t dyn = 0;
#define BWD_MASKS
#endif
    if (dL_dout_aux)
        launch_chain(render_backward_kernel<true>, grid, dim3(BWD_THREADS), dyn, s, v, g.rec0, g.rec1, g.rec2, im.starts,
                     b.points, BWD_MASKS im.final_T, im.n_contrib, dL_dout, dL_dout_aux, scratch);
    else
        launch_chain(render_backward_kernel<false>, grid, dim3(BWD_THREADS), dyn, s, v, g.rec0, g.rec1, g.rec2, im.starts,
                     b.points, BWD_MASKS im.final_T, im.n_contrib, dL_dout, (const float*)nullptr, scratch);
}

}  // namespace ggrt
