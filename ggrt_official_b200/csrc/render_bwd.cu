// Per-tile back-to-front gradient replay (A.4).  Replaces upstream renderCUDA (bwd),
// SURVEY.md 8a row a12 -- the dominant cost of fwd+bwd upstream because every pixel
// issues ~10 global float atomics per contributing Gaussian.
//
// Work decomposition ("Gaussian-parallel"): one CTA per 16x16 tile, warp w owns the 8x4
// pixel block (w&1, w>>1) as in the forward.  A batch of up to 512 Gaussian records is
// staged into shared memory; each warp culls it against its pixel block (exact ellipse vs
// rectangle test on {alpha >= 1/255}) into a back-to-front queue and then consumes the queue
// GL = 8 Gaussians at a time with LANES = (pixel slot, Gaussian slot): per step the warp
// evaluates 8 Gaussians at 4 pixels, a segmented warp scan over the associative maps
// (T, R) -> (aT, R + bT) recovers each Gaussian's transmittance T_i and the colour
// accumulated behind it, and every lane accumulates the 9 partial gradients of ITS Gaussian
// in registers.  There is no per-Gaussian reduction over 32 pixel lanes (the upstream
// bottleneck): the 4 pixel-slot partials are folded once per chunk and committed with
// vector RED.ADD.F32 (2 x v4 + 1 scalar per Gaussian per warp).
//   dL/dalpha_i = T_i (c_i . g) - [ sum_{j behind i} w_j (c_j . g) + T_final (bg . g) ] / (1 - alpha_i)
// Per-pixel running state {T behind, sum behind} lives in shared memory between chunks.
#include "render_common.cuh"

namespace ggrt {

// EXPERIMENTAL, off by default: two 4x4 half-block queues per warp instead of
// one 8x4 queue -- the CPU step model predicts -25 % pixel steps, +49 % chunks; measured once at C2: 174.0 -> 167.5 us,
// parity + edge-case tests green.  To become the default it still needs the full GPU suite.
#ifndef GGRT_BWD_HALVES
#define GGRT_BWD_HALVES 0
#endif
#ifndef GGRT_BWD_BATCH
#define GGRT_BWD_BATCH (GGRT_BWD_HALVES ? 384 : 512)  // the half queues need shared memory: keep 4 CTAs per SM
#endif
#ifndef GGRT_BWD_MINBLOCKS
#define GGRT_BWD_MINBLOCKS 4
#endif
constexpr int BWD_BATCH = GGRT_BWD_BATCH;
constexpr int NWARPS = BWD_WARPS;
#ifndef GGRT_BWD_PIX_UNROLL
#define GGRT_BWD_PIX_UNROLL 2
#endif
constexpr int PIX_UNROLL = GGRT_BWD_PIX_UNROLL;
#ifndef GGRT_BWD_GL
#define GGRT_BWD_GL 8
#endif
constexpr int GL = GGRT_BWD_GL;   // Gaussians per warp step (power of two <= 32)
constexpr int PL = 32 / GL;       // pixels per warp step
constexpr int NHALF = GGRT_BWD_HALVES ? 2 : 1;  // queues per warp
static_assert(!GGRT_BWD_HALVES || PL == 4, "half-block queues assume 4-pixel groups (one half-row each)");

__device__ __forceinline__ void red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// 16-byte vector reduction (sm_90+): one L2 operation for four adjacent floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// AUX: the 4th blended channel (out_depth) also carries an upstream gradient.
template <bool AUX>
__global__ void __launch_bounds__(BWD_THREADS, GGRT_BWD_MINBLOCKS * 8 / BWD_WARPS)
render_backward_kernel(View v, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                       const float4* __restrict__ rec2, const uint32_t* __restrict__ starts,
                       const uint32_t* __restrict__ points, const float* __restrict__ final_T,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dout,
                       const float* __restrict__ dL_dout_aux, float* __restrict__ scratch) {
    __shared__ __align__(16) unsigned char srec[BWD_BATCH * REC_BYTES];
    __shared__ uint32_t sid[BWD_BATCH];
    __shared__ unsigned short squeue[NWARPS][BWD_BATCH];
    __shared__ unsigned short shalf[GGRT_BWD_HALVES ? NWARPS : 1][2][GGRT_BWD_HALVES ? BWD_BATCH : 1];
    __shared__ float4 spix_g[NWARPS][32];   // per pixel {g_r, g_g, g_b, bits(last)}
    __shared__ float4 spix_s[NWARPS][32];   // per pixel {x, y, T behind, sum behind}
    __shared__ float spix_a[AUX ? NWARPS : 1][32];  // per pixel gradient of the aux channel
    __shared__ uint32_t block_last_s;

    const uint32_t sbase = smem_addr(srec);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int wt = warp + blockIdx.z * BWD_WARPS;  // warp pixel block of the tile (8 per tile)
    const int bx0 = blockIdx.x * TILE + (wt & 1) * 8, by0 = blockIdx.y * TILE + (wt >> 1) * 4;
    const float bx0f = (float)bx0, by0f = (float)by0;
    const uint32_t start = starts[tile];

    // ---- per-pixel constants / initial state (lane = pixel here) -------------------------------
    uint32_t last = 0;
    {
        const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
        float Tfin = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, da = 0.f;
        if (px < v.W && py < v.H) {
            const size_t pix = (size_t)py * v.W + px, hw = (size_t)v.H * v.W;
            Tfin = final_T[pix];
            last = n_contrib[pix];
            d0 = dL_dout[pix];
            d1 = dL_dout[hw + pix];
            d2 = dL_dout[2 * hw + pix];
            if (AUX) da = dL_dout_aux[pix];
        }
        // pixels with a zero upstream gradient contribute nothing (crop training leaves most tiles empty)
        if (d0 == 0.f && d1 == 0.f && d2 == 0.f && da == 0.f) last = 0;
        if (AUX) spix_a[warp][lane] = da;
        const float bg_dot = v.bg[0] * d0 + v.bg[1] * d1 + v.bg[2] * d2;
        spix_g[warp][lane] = make_float4(d0, d1, d2, __uint_as_float(last));
        spix_s[warp][lane] = make_float4((float)px, (float)py, Tfin, Tfin * bg_dot);
    }
    if (tid == 0) block_last_s = 0;
    __syncthreads();
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0 && warp_last > 0) atomicMax(&block_last_s, warp_last);
    __syncthreads();
    const uint32_t block_last = block_last_s;
    if (block_last == 0) return;
    const uint32_t pmask = __ballot_sync(0xffffffffu, last > 0);  // pixels of this warp that matter

    const float neg_half_w = -0.5f * (float)v.W, neg_half_h = -0.5f * (float)v.H;
    const uint32_t gaddr = smem_addr(&spix_g[warp][0]), saddr = smem_addr(&spix_s[warp][0]);
    const float* aux_g = &spix_a[AUX ? warp : 0][0];

    const int nb = (int)((block_last + BWD_BATCH - 1) / BWD_BATCH);
    for (int bi = nb - 1; bi >= 0; --bi) {
        const uint32_t boff = (uint32_t)bi * BWD_BATCH;
        const uint32_t cnt = min((uint32_t)BWD_BATCH, block_last - boff);
        __syncthreads();  // every warp is done with the previous batch before the refill
        for (uint32_t k = tid; k < cnt; k += BWD_THREADS) {
            const uint32_t id = points[start + boff + k];
            sid[k] = id;
            const uint32_t dst = sbase + k * REC_BYTES;
            sts128(dst, rec0[id]);
            sts128(dst + 16, rec1[id]);
            sts128(dst + 32, rec2[id]);
        }
        __syncthreads();
        if (warp_last <= boff) continue;

        // ---- cull the batch against this warp's pixel block into a back-to-front queue -----------
        uint32_t qn = 0;
        const uint32_t lim = min(cnt, warp_last - boff);  // entries at or beyond warp_last never contribute here
        for (int r = (int)((lim - 1) & ~31u); r >= 0; r -= 32) {
            const uint32_t j = (uint32_t)r + 31 - lane;  // lane 0 tests the backmost entry of the round
            bool hit = false;
            if (j < lim) {
                const float4 a = lds128(sbase + j * REC_BYTES);
                const float4 c = lds128(sbase + j * REC_BYTES + 16);
                hit = ellipse_hits_rect(a.x, a.y, a.z, c.x, c.y, c.z, bx0f, by0f, 7.0f, 3.0f);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) squeue[warp][qn + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;
            qn += __popc(m);
        }
        __syncwarp();

        // ---- GL queued Gaussians at a time.  Lane tiling: lane = (pixel slot q, Gaussian slot gl); the warp
        // processes GL Gaussians x PL = 32/GL pixels per step, so the scan has log2(GL) levels and the per-pixel
        // work is shared by fewer idle lanes.  The PL partial accumulators of a Gaussian are combined at the end.
        const int gl = lane & (GL - 1), q = lane / GL;
        uint32_t qh0 = 0, qh1 = 0;
        if (GGRT_BWD_HALVES) {
            // second-level cull, over the survivors only: which of the two 4x4 half blocks does each queued Gaussian
            // reach?  Order (back to front) is preserved.  A Gaussian that misses a half is inactive at every pixel of
            // it, i.e. the identity of the (T, R) recurrence there, so leaving it out of that half's queue is exact.
            for (uint32_t e0 = 0; e0 < qn; e0 += 32) {
                const uint32_t e = e0 + lane;
                bool hit_l = false, hit_r = false;
                unsigned short j = 0;
                if (e < qn) {
                    j = squeue[warp][e];
                    const float4 a = lds128(sbase + (uint32_t)j * REC_BYTES);
                    const float4 c = lds128(sbase + (uint32_t)j * REC_BYTES + 16);
                    hit_l = ellipse_hits_rect(a.x, a.y, a.z, c.x, c.y, c.z, bx0f, by0f, 3.0f, 3.0f);
                    hit_r = ellipse_hits_rect(a.x, a.y, a.z, c.x, c.y, c.z, bx0f + 4.0f, by0f, 3.0f, 3.0f);
                }
                const uint32_t ml = __ballot_sync(0xffffffffu, hit_l), mr = __ballot_sync(0xffffffffu, hit_r);
                const uint32_t below = (1u << lane) - 1u;
                if (hit_l) shalf[GGRT_BWD_HALVES ? warp : 0][0][qh0 + __popc(ml & below)] = j;
                if (hit_r) shalf[GGRT_BWD_HALVES ? warp : 0][1][qh1 + __popc(mr & below)] = j;
                qh0 += __popc(ml), qh1 += __popc(mr);
            }
            __syncwarp();
        }
        for (int half = 0; half < NHALF; ++half) {
        const unsigned short* queue = GGRT_BWD_HALVES ? shalf[GGRT_BWD_HALVES ? warp : 0][half] : squeue[warp];
        const uint32_t qcount = GGRT_BWD_HALVES ? (half == 0 ? qh0 : qh1) : qn;
        for (uint32_t c0 = 0; c0 < qcount; c0 += GL) {
            const bool valid = c0 + gl < qcount;
            const uint32_t jj = valid ? queue[c0 + gl] : 0u;
            const uint32_t src = sbase + jj * REC_BYTES;
            const float2 gxy = lds64(src);
            float4 con = lds128(src + 16);
            const float4 col = lds128(src + 32);
            if (!valid) con.w = 0.f;  // zero opacity: never active
            uint32_t pos = valid ? boff + jj : 0xffffffffu;  // 0-based list position (never "behind" a pixel's last)
            asm volatile("" : "+r"(pos));  // keep the select: otherwise `valid` is re-tested in every pixel step
            // exponent of the Gaussian in base 2 with the -1/2 folded in: G = 2^(ea dx^2 + eb dx dy + ec dy^2)
            const float ea = -0.5f * LOG2E * con.x, eb = -LOG2E * con.y, ec = -0.5f * LOG2E * con.z;
            float a_op = 0.f, a_mx = 0.f, a_my = 0.f, a_A = 0.f, a_B = 0.f, a_C = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f;
            float a_x = 0.f;  // aux channel

            // pixel groups: group t holds pixels t*PL .. t*PL+PL-1, one per pixel slot (with half-block queues the
            // even groups are the left 4x4 half, the odd ones the right half)
#pragma unroll PIX_UNROLL
            for (int tt = 0; tt < 32 / PL / NHALF; ++tt) {
                const int t = tt * NHALF + half;
                if (((pmask >> (t * PL)) & ((1u << PL) - 1u)) == 0) continue;  // no pixel of the group matters
                const int p = t * PL + q;
                const float4 pg = lds128(gaddr + p * 16);   // {g_r, g_g, g_b, last}
                const float4 ps = lds128(saddr + p * 16);   // {x, y, T behind, sum behind}
                const float dx = gxy.x - ps.x, dy = gxy.y - ps.y;
                const float dxx = dx * dx, dyy = dy * dy, dxy = dx * dy;
                const float power2 = fmaf(ea, dxx, fmaf(ec, dyy, eb * dxy));  // log2 of the Gaussian
                const float G = ex2_approx(power2);
                const float alpha_raw = fminf(ALPHA_MAX, con.w * G);
                const bool act = (pos < __float_as_uint(pg.w)) && (power2 <= 0.0f) && (alpha_raw >= ALPHA_MIN);
                const float alpha = act ? alpha_raw : 0.f;
                const float inv_om = rcp_approx(1.0f - alpha);
                float sdot = fmaf(col.z, pg.z, fmaf(col.y, pg.y, col.x * pg.x));
                float ga = 0.f;
                if (AUX) {
                    ga = aux_g[p];
                    sdot = fmaf(col.w, ga, sdot);
                }
                // Going back to front each Gaussian maps the running pair (T, R) to (a T, R + b T) with
                // a = 1/(1-alpha), b = alpha (c.g)/(1-alpha).  These maps compose associatively, so ONE
                // segmented warp scan (slot 0 = backmost) yields every lane's transmittance and the sum behind it.
                float A = inv_om, B = alpha * sdot * inv_om;
#pragma unroll
                for (int d = 1; d < GL; d <<= 1) {
                    const float Ap = __shfl_up_sync(0xffffffffu, A, d, GL);
                    const float Bp = __shfl_up_sync(0xffffffffu, B, d, GL);
                    if (gl >= d) {
                        B = fmaf(B, Ap, Bp);
                        A *= Ap;
                    }
                }
                const float Ti = ps.z * A;                   // transmittance in front of this Gaussian
                const float w = alpha * Ti;
                const float Rtot = fmaf(ps.z, B, ps.w);      // sum behind, this Gaussian included
                const float dL_dalpha = fmaf(Ti, sdot, -(Rtot - w * sdot) * inv_om);
                const float qv = act ? G * dL_dalpha : 0.f;
                const float tq = con.w * qv;
                a_op += qv;
                a_mx = fmaf(tq, dx, a_mx);  // first moments; the conic is applied once per chunk below
                a_my = fmaf(tq, dy, a_my);
                a_A = fmaf(tq, dxx, a_A);
                a_B = fmaf(tq, dxy, a_B);
                a_C = fmaf(tq, dyy, a_C);
                a_r = fmaf(w, pg.x, a_r);
                a_g = fmaf(w, pg.y, a_g);
                a_b = fmaf(w, pg.z, a_b);
                if (AUX) a_x = fmaf(w, ga, a_x);
                if (gl == GL - 1) {  // frontmost slot holds the chunk totals: state behind the next chunk
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr + p * 16 + 8), "f"(Ti), "f"(Rtot)
                                 : "memory");
                }
            }
            __syncwarp();
            // combine the PL pixel-slot partials of each Gaussian (lanes gl, gl+GL, ...)
#pragma unroll
            for (int d = GL; d < 32; d <<= 1) {
                a_op += __shfl_xor_sync(0xffffffffu, a_op, d);
                a_mx += __shfl_xor_sync(0xffffffffu, a_mx, d);
                a_my += __shfl_xor_sync(0xffffffffu, a_my, d);
                a_A += __shfl_xor_sync(0xffffffffu, a_A, d);
                a_B += __shfl_xor_sync(0xffffffffu, a_B, d);
                a_C += __shfl_xor_sync(0xffffffffu, a_C, d);
                a_r += __shfl_xor_sync(0xffffffffu, a_r, d);
                a_g += __shfl_xor_sync(0xffffffffu, a_g, d);
                a_b += __shfl_xor_sync(0xffffffffu, a_b, d);
                if (AUX) a_x += __shfl_xor_sync(0xffffffffu, a_x, d);
            }
            if (valid && q == 0) {
                float* dst = scratch + (size_t)sid[jj] * GRAD_STRIDE;
                // scratch rows are 48 B (16-B aligned): slots {mx,my,A,B} {C,op,r,g} {b}
                const float gmx = fmaf(con.x, a_mx, con.y * a_my), gmy = fmaf(con.z, a_my, con.y * a_mx);
                red_add_v4(dst + G_MX, gmx * neg_half_w, gmy * neg_half_h, -0.5f * a_A, -0.5f * a_B);
                red_add_v4(dst + G_CC, -0.5f * a_C, a_op, a_r, a_g);
                red_add(dst + G_B, a_b);
                if (AUX) red_add(dst + G_AUX, a_x);
            }
        }
        }  // half
    }
}

void launch_render_backward(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, const float* dL_dout,
                            const float* dL_dout_aux, float* scratch, cudaStream_t s) {
    dim3 grid(v.gx, v.gy, 8 / BWD_WARPS);
    if (dL_dout_aux)
        render_backward_kernel<true><<<grid, BWD_THREADS, 0, s>>>(v, g.rec0, g.rec1, g.rec2, im.starts, b.points,
                                                                     im.final_T, im.n_contrib, dL_dout, dL_dout_aux,
                                                                     scratch);
    else
        render_backward_kernel<false><<<grid, BWD_THREADS, 0, s>>>(v, g.rec0, g.rec1, g.rec2, im.starts, b.points,
                                                                      im.final_T, im.n_contrib, dL_dout, nullptr, scratch);
}

}  // namespace ggrt
