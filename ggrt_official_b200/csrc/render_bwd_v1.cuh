// Round-2 first version of the render backward (kept for A/B builds: -DGGRT_BWD_MMA=0): 8 Gaussians x 8 pixels per
// warp step, per-lane register accumulators, shuffle fold per chunk.  See render_bwd.cu for the shared helpers.
#pragma once

namespace ggrt {

// Per-warp pixel tables, indexed by pair slot j = row * 4 + q (pixels (q, row) = "A" and (q + 4, row) = "B"):
//   spix[j] = {g_r A, g_r B, g_g A, g_g B | g_b A, g_b B, last A (bits), last B (bits) |
//              T behind A, T behind B, -(sum behind) A, -(sum behind) B}          sga[j] = {g_aux A, g_aux B}
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
    __shared__ __align__(16) float spix[NWARPS][16][12];  // per pair slot: sg0 | sg1 | sst (48 B, one base register)
    __shared__ __align__(8) float sga[AUX ? NWARPS : 1][16][2];
    __shared__ uint32_t block_last_s;

    const uint32_t sbase = smem_addr(srec);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int wt = warp + blockIdx.z * BWD_WARPS;  // warp pixel block of the tile (8 per tile)
    const int bx0 = blockIdx.x * TILE + (wt & 1) * 8, by0 = blockIdx.y * TILE + (wt >> 1) * 4;
    const float bx0f = (float)bx0, by0f = (float)by0;
    pdl_enter();
    const uint32_t start = starts[tile];

    // ---- per-pixel constants / initial state (lane = pixel here) -------------------------------
    uint32_t last = 0;
    {
        const int lx = lane & 7, ly = lane >> 3;
        const int px = bx0 + lx, py = by0 + ly;
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
        const float bg_dot = v.bg[0] * d0 + v.bg[1] * d1 + v.bg[2] * d2;
        const int j = ly * 4 + (lx & 3), h = lx >> 2;  // pair slot, half (0 = A, 1 = B)
        float* sp = spix[warp][j];
        sp[h] = d0, sp[2 + h] = d1;
        sp[4 + h] = d2, sp[6 + h] = __uint_as_float(last);
        sp[8 + h] = Tfin, sp[10 + h] = -(Tfin * bg_dot);
        if (AUX) sga[warp][j][h] = da;
    }
    if (tid == 0) block_last_s = 0;
    __syncthreads();
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0 && warp_last > 0) atomicMax(&block_last_s, warp_last);
    __syncthreads();
    const uint32_t block_last = block_last_s;
    if (block_last == 0) return;
    const uint32_t pmask = __ballot_sync(0xffffffffu, last > 0);  // pixels of this warp that matter (bit = ly*8 + lx)

    const float neg_half_w = -0.5f * (float)v.W, neg_half_h = -0.5f * (float)v.H;
    const int gl = lane & (GL - 1), q = lane / GL;
    const uint32_t paddr = smem_addr(&spix[warp][q][0]);  // + row * 192 (4 pair slots of 48 B per row)
    const float* aux_g = &sga[AUX ? warp : 0][q][0];
    const float xq = bx0f + (float)q;  // x of this lane's pixel A; pixel B is 4 to the right

    const int nb = (int)((block_last + BWD_BATCH - 1) / BWD_BATCH);
    for (int bi = nb - 1; bi >= 0; --bi) {
        const uint32_t boff = (uint32_t)bi * BWD_BATCH;
        const uint32_t cnt = min((uint32_t)BWD_BATCH, block_last - boff);
        __syncthreads();  // every warp is done with the previous batch before the refill
        for (uint32_t k = tid; k < cnt; k += BWD_THREADS) {  // gather the batch's records with cp.async (LDGSTS)
            const uint32_t id = points[start + boff + k];
            sid[k] = id;
            const uint32_t dst = sbase + k * REC_BYTES;
            cp_async16(dst, rec0 + id);
            cp_async16(dst + 16, rec1 + id);
            cp_async16(dst + 32, rec2 + id);
        }
        cp_async_commit();
        cp_async_wait<0>();
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

        // ---- 8 queued Gaussians at a time; lane = (pixel slot q, Gaussian slot gl) ------------------
        const unsigned short* queue = squeue[warp];
        for (uint32_t c0 = 0; c0 < qn; c0 += GL) {
            const bool valid = c0 + gl < qn;
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
            const float ncw = -con.w;
            // per-chunk constants of this lane's pixel pair: dx (B is 4 px to the right of A), dx^2, ea dx^2
            const float dxa = gxy.x - xq;
            const f2 dx2 = pk(dxa, dxa - 4.0f);
            const f2 pa2 = mul2(bc(ea), mul2(dx2, dx2));
            const float gyr = gxy.y - by0f;
            f2 a_op2 = bc(0.f), a_mx2 = bc(0.f), a_A2 = bc(0.f), a_B2 = bc(0.f);
            f2 a_r2 = bc(0.f), a_g2 = bc(0.f), a_b2 = bc(0.f), a_x2 = bc(0.f);  // colour sums come out NEGATED
            float a_my = 0.f, a_C = 0.f;

#pragma unroll
            for (int row = 0; row < 4; ++row) {
                if (((pmask >> (row * 8)) & 0xffu) == 0) continue;  // no pixel of the row matters
                const uint32_t off = paddr + (uint32_t)row * 192u;  // an immediate offset once the loop is unrolled
                f2 gr2, gg2, gb2, last2, Tb2, nRb2;
                lds_2f2(off, gr2, gg2);
                lds_2f2(off + 16, gb2, last2);
                lds_2f2(off + 32, Tb2, nRb2);
                const float dy = gyr - (float)row;
                // log2 of the Gaussian at both pixels: ea dx^2 + (eb dx + ec dy) dy
                const f2 pw2 = fma2(fma2(bc(eb), dx2, bc(ec * dy)), bc(dy), pa2);
                const float pwa = lo(pw2), pwb = hi(pw2);
                const float Ga = ex2_approx(pwa), Gb = ex2_approx(pwb);
                const f2 G2 = pk(Ga, Gb);
                const f2 arn2 = mul2(bc(ncw), G2);  // -(opacity * G)
                const float arna = fmaxf(-ALPHA_MAX, lo(arn2)), arnb = fmaxf(-ALPHA_MAX, hi(arn2));
                const bool acta = (pos < __float_as_uint(lo(last2))) && (pwa <= 0.0f) && (arna <= -ALPHA_MIN);
                const bool actb = (pos < __float_as_uint(hi(last2))) && (pwb <= 0.0f) && (arnb <= -ALPHA_MIN);
                const f2 nal2 = pk(acta ? arna : 0.f, actb ? arnb : 0.f);  // -alpha
                const f2 om2 = add2(bc(1.0f), nal2);
                const f2 io2 = pk(rcp_approx(lo(om2)), rcp_approx(hi(om2)));  // 1 / (1 - alpha)
                f2 sdot2 = fma2(bc(col.z), gb2, fma2(bc(col.y), gg2, mul2(bc(col.x), gr2)));
                f2 ga2 = bc(0.f);
                if (AUX) {
                    ga2 = lds_f2(smem_addr(aux_g) + (uint32_t)row * 32u);
                    sdot2 = fma2(bc(col.w), ga2, sdot2);
                }
                // Going back to front each Gaussian maps the running pair (T, -R) to (a T, -R + nb T) with
                // a = 1/(1-alpha), nb = -alpha (c.g)/(1-alpha).  These maps compose associatively, so ONE
                // segmented warp scan (slot 0 = backmost) yields every lane's transmittance and the sum behind it.
                f2 A2 = io2, nB2 = mul2(mul2(nal2, sdot2), io2);
#pragma unroll
                for (int d = 1; d < GL; d <<= 1) scan_step(A2, nB2, d, gl);
                const f2 Ti2 = mul2(Tb2, A2);            // transmittance in front of this Gaussian
                const f2 nRt2 = fma2(Tb2, nB2, nRb2);    // -(sum behind, this Gaussian included)
                const f2 dLa2 = mul2(fma2(Ti2, sdot2, nRt2), io2);  // dL/dalpha
                const f2 qg2 = mul2(G2, dLa2);
                const f2 qv2 = pk(acta ? lo(qg2) : 0.f, actb ? hi(qg2) : 0.f);
                const f2 tq2 = mul2(bc(con.w), qv2);
                const f2 nw2 = mul2(nal2, Ti2);          // -(blend weight)
                a_op2 = add2(a_op2, qv2);
                const f2 inc2 = mul2(tq2, dx2);          // first x moment; the conic is applied once per chunk below
                a_mx2 = add2(a_mx2, inc2);
                a_B2 = fma2(inc2, bc(dy), a_B2);
                a_A2 = fma2(inc2, dx2, a_A2);
                const float sd = (lo(tq2) + hi(tq2)) * dy;  // the pair shares dy: y moments once for both pixels
                a_my += sd;
                a_C = fmaf(sd, dy, a_C);
                a_r2 = fma2(nw2, gr2, a_r2);
                a_g2 = fma2(nw2, gg2, a_g2);
                a_b2 = fma2(nw2, gb2, a_b2);
                if (AUX) a_x2 = fma2(nw2, ga2, a_x2);
                if (gl == GL - 1) sts_2f2(off + 32, Ti2, nRt2);  // frontmost slot: state behind the next chunk
            }
            __syncwarp();
            // fold the pair, then the QL pixel-slot partials of each Gaussian (lanes gl, gl+GL, ...)
            float a_op = lo(a_op2) + hi(a_op2), a_mx = lo(a_mx2) + hi(a_mx2), a_A = lo(a_A2) + hi(a_A2),
                  a_B = lo(a_B2) + hi(a_B2), a_r = lo(a_r2) + hi(a_r2), a_g = lo(a_g2) + hi(a_g2),
                  a_b = lo(a_b2) + hi(a_b2), a_x = lo(a_x2) + hi(a_x2);
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
                red_add_v4(dst + G_CC, -0.5f * a_C, a_op, -a_r, -a_g);
                red_add(dst + G_B, -a_b);
                if (AUX) red_add(dst + G_AUX, -a_x);
            }
        }
    }
}

}  // namespace ggrt
