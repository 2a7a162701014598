// Per-tile back-to-front gradient replay (A.4).  Replaces upstream renderCUDA (bwd),
// SURVEY.md 8a row a12 -- the dominant cost of fwd+bwd upstream because every pixel
// issues ~10 global float atomics per contributing Gaussian.
//
// Work decomposition ("Gaussian-parallel"): one CTA per 16x16 tile, warp w owns the 8x4
// pixel block (w&1, w>>1) as in the forward.  A batch of up to 512 Gaussian records is
// staged into shared memory; each warp culls it against its pixel block (bounding box of
// {alpha >= 1/255}) into a back-to-front queue and then processes the queue 32 Gaussians
// at a time with LANES = GAUSSIANS: for every pixel of the block the 32 lanes evaluate
// their Gaussian's alpha, a warp prefix product of (1-alpha) recovers each Gaussian's
// transmittance T_i from the transmittance behind the chunk, and a warp prefix sum gives
// the colour accumulated behind it.  Every lane accumulates the 9 partial gradients of ITS
// Gaussian over the block's pixels in registers, so there is no per-Gaussian cross-lane
// reduction (the upstream bottleneck) -- one set of global RED.ADD.F32 per lane per chunk.
//   dL/dalpha_i = T_i (c_i . g) - [ sum_{j behind i} w_j (c_j . g) + T_final (bg . g) ] / (1 - alpha_i)
// Per-pixel running state {T behind, sum behind} lives in shared memory between chunks.
#include "render_common.cuh"

namespace ggrt {

constexpr int BWD_BATCH = 512;
constexpr int NWARPS = RENDER_THREADS / 32;
#ifndef GGRT_BWD_PIX_UNROLL
#define GGRT_BWD_PIX_UNROLL 2
#endif
constexpr int PIX_UNROLL = GGRT_BWD_PIX_UNROLL;

__device__ __forceinline__ void red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// 16-byte vector reduction (sm_90+): one L2 operation for four adjacent floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// AUX: the 4th blended channel (out_depth) also carries an upstream gradient.
template <bool AUX>
__global__ void __launch_bounds__(RENDER_THREADS, 3)
render_backward_kernel(View v, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                       const float4* __restrict__ rec2, const uint32_t* __restrict__ starts,
                       const uint32_t* __restrict__ points, const float* __restrict__ final_T,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dout,
                       const float* __restrict__ dL_dout_aux, float* __restrict__ scratch) {
    __shared__ __align__(16) unsigned char srec[BWD_BATCH * REC_BYTES];
    __shared__ uint32_t sid[BWD_BATCH];
    __shared__ unsigned short squeue[NWARPS][BWD_BATCH];
    __shared__ float4 spix_g[NWARPS][32];   // per pixel {g_r, g_g, g_b, bits(last)}
    __shared__ float4 spix_s[NWARPS][32];   // per pixel {x, y, T behind, sum behind}
    __shared__ float spix_a[AUX ? NWARPS : 1][32];  // per pixel gradient of the aux channel
    __shared__ uint32_t block_last_s;

    const uint32_t sbase = smem_addr(srec);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int bx0 = blockIdx.x * TILE + (warp & 1) * 8, by0 = blockIdx.y * TILE + (warp >> 1) * 4;
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
        for (uint32_t k = tid; k < cnt; k += RENDER_THREADS) {
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

        // ---- 32 queued Gaussians at a time: lane = Gaussian ----------------------------------------
        for (uint32_t c0 = 0; c0 < qn; c0 += 32) {
            const bool valid = c0 + lane < qn;
            const uint32_t jj = valid ? squeue[warp][c0 + lane] : 0u;
            const uint32_t src = sbase + jj * REC_BYTES;
            const float2 gxy = lds64(src);
            float4 con = lds128(src + 16);
            const float4 col = lds128(src + 32);
            if (!valid) con.w = 0.f;  // zero opacity: never active
            const uint32_t pos = valid ? boff + jj : 0xffffffffu;  // 0-based list position
            // exponent of the Gaussian in base 2 with the -1/2 folded in: G = 2^(ea dx^2 + eb dx dy + ec dy^2)
            const float ea = -0.5f * LOG2E * con.x, eb = -LOG2E * con.y, ec = -0.5f * LOG2E * con.z;
            float a_op = 0.f, a_mx = 0.f, a_my = 0.f, a_A = 0.f, a_B = 0.f, a_C = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f;
            float a_x = 0.f;  // aux channel

            // PIX_UNROLL pixels are processed per iteration: their scans are independent dependency chains
            // that the scheduler interleaves (the kernel is shuffle-latency bound otherwise).
            uint32_t pm = pmask;
            while (pm) {
                int p[PIX_UNROLL];
                bool pv[PIX_UNROLL];
#pragma unroll
                for (int u = 0; u < PIX_UNROLL; ++u) {
                    pv[u] = pm != 0;
                    p[u] = pv[u] ? __ffs(pm) - 1 : p[0];
                    pm &= pm - 1;
                }
                float4 pg[PIX_UNROLL], ps[PIX_UNROLL];
                float dx[PIX_UNROLL], dy[PIX_UNROLL], dxx[PIX_UNROLL], dyy[PIX_UNROLL], dxy[PIX_UNROLL];
                float G[PIX_UNROLL], alpha[PIX_UNROLL], inv_om[PIX_UNROLL], s[PIX_UNROLL], A[PIX_UNROLL], B[PIX_UNROLL];
                bool act[PIX_UNROLL];
#pragma unroll
                for (int u = 0; u < PIX_UNROLL; ++u) {
                    pg[u] = lds128(gaddr + p[u] * 16);   // {g_r, g_g, g_b, last}
                    ps[u] = lds128(saddr + p[u] * 16);   // {x, y, T behind, sum behind}
                }
#pragma unroll
                for (int u = 0; u < PIX_UNROLL; ++u) {
                    dx[u] = gxy.x - ps[u].x, dy[u] = gxy.y - ps[u].y;
                    dxx[u] = dx[u] * dx[u], dyy[u] = dy[u] * dy[u], dxy[u] = dx[u] * dy[u];
                    const float power2 = fmaf(ea, dxx[u], fmaf(ec, dyy[u], eb * dxy[u]));  // log2 of the Gaussian
                    G[u] = ex2_approx(power2);
                    const float alpha_raw = fminf(ALPHA_MAX, con.w * G[u]);
                    act[u] = pv[u] && (pos < __float_as_uint(pg[u].w)) && (power2 <= 0.0f) && (alpha_raw >= ALPHA_MIN);
                    alpha[u] = act[u] ? alpha_raw : 0.f;
                    inv_om[u] = rcp_approx(1.0f - alpha[u]);
                    s[u] = fmaf(col.z, pg[u].z, fmaf(col.y, pg[u].y, col.x * pg[u].x));
                    if (AUX) s[u] = fmaf(col.w, aux_g[p[u]], s[u]);
                    // Going back to front each Gaussian maps the running pair (T, R) to (a T, R + b T) with
                    // a = 1/(1-alpha), b = alpha (c.g)/(1-alpha).  These maps compose associatively, so ONE
                    // warp scan (lane 0 = backmost) yields every lane's transmittance and the sum behind it.
                    A[u] = inv_om[u];
                    B[u] = alpha[u] * s[u] * inv_om[u];
                }
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                    for (int u = 0; u < PIX_UNROLL; ++u) {
                        const float Ap = __shfl_up_sync(0xffffffffu, A[u], d);
                        const float Bp = __shfl_up_sync(0xffffffffu, B[u], d);
                        if (lane >= d) {
                            B[u] = fmaf(B[u], Ap, Bp);
                            A[u] *= Ap;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < PIX_UNROLL; ++u) {
                    const float Ti = ps[u].z * A[u];                    // transmittance in front of this Gaussian
                    const float w = alpha[u] * Ti;
                    const float Rtot = fmaf(ps[u].z, B[u], ps[u].w);   // sum behind, this Gaussian included
                    const float dL_dalpha = fmaf(Ti, s[u], -(Rtot - w * s[u]) * inv_om[u]);
                    const float q = act[u] ? G[u] * dL_dalpha : 0.f;
                    const float t = con.w * q;
                    a_op += q;
                    a_mx = fmaf(t, dx[u], a_mx);  // first moments; the conic is applied once per chunk below
                    a_my = fmaf(t, dy[u], a_my);
                    a_A = fmaf(t, dxx[u], a_A);
                    a_B = fmaf(t, dxy[u], a_B);
                    a_C = fmaf(t, dyy[u], a_C);
                    a_r = fmaf(w, pg[u].x, a_r);
                    a_g = fmaf(w, pg[u].y, a_g);
                    a_b = fmaf(w, pg[u].z, a_b);
                    if (AUX) a_x = fmaf(w, aux_g[p[u]], a_x);
                    if (lane == 31 && pv[u]) {  // frontmost lane holds the chunk totals: state behind the next chunk
                        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr + p[u] * 16 + 8), "f"(Ti), "f"(Rtot)
                                     : "memory");
                    }
                }
            }
            __syncwarp();
            if (valid) {
                float* dst = scratch + (size_t)sid[jj] * GRAD_STRIDE;
                // scratch rows are 48 B (16-B aligned): slots {mx,my,A,B} {C,op,r,g} {b}
                const float gmx = fmaf(con.x, a_mx, con.y * a_my), gmy = fmaf(con.z, a_my, con.y * a_mx);
                red_add_v4(dst + G_MX, gmx * neg_half_w, gmy * neg_half_h, -0.5f * a_A, -0.5f * a_B);
                red_add_v4(dst + G_CC, -0.5f * a_C, a_op, a_r, a_g);
                red_add(dst + G_B, a_b);
                if (AUX) red_add(dst + G_AUX, a_x);
            }
        }
    }
}

void launch_render_backward(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, const float* dL_dout,
                            const float* dL_dout_aux, float* scratch, cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    if (dL_dout_aux)
        render_backward_kernel<true><<<grid, RENDER_THREADS, 0, s>>>(v, g.rec0, g.rec1, g.rec2, im.starts, b.points,
                                                                     im.final_T, im.n_contrib, dL_dout, dL_dout_aux,
                                                                     scratch);
    else
        render_backward_kernel<false><<<grid, RENDER_THREADS, 0, s>>>(v, g.rec0, g.rec1, g.rec2, im.starts, b.points,
                                                                      im.final_T, im.n_contrib, dL_dout, nullptr, scratch);
}

}  // namespace ggrt
