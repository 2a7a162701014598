// Per-tile back-to-front gradient replay (A.4).  Replaces upstream renderCUDA (bwd),
// SURVEY.md 8a row a12 -- the dominant cost of fwd+bwd upstream because every pixel
// issues ~10 global float atomics per contributing Gaussian.
//
// Same tile / warp / batch structure and the same per-warp bounding-box cull as the
// forward.  Per (warp, Gaussian) the 9 partial gradients are reduced over the 32 pixels
// with a transposed shuffle butterfly (14 shuffles for 9 values, leaving value k in lane
// 4k) and committed with two warp-level global reductions (RED.ADD.F32: 8 lanes hit one
// 48-byte scratch row, then the 9th value).  No shared-memory float atomics: those
// compile to compare-and-swap loops on this architecture.
#include "render_common.cuh"

namespace ggrt {

constexpr int NV = 9;  // gradient values per Gaussian (GradSlot)

__device__ __forceinline__ void red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}

// Transposed butterfly: reduces g[0..7] over the warp with 4+2+1+1+1 = 9 shuffles instead of
// 8*5 = 40 (each step halves the number of live values per lane).  On return every lane l
// holds the warp total of value (l >> 2).
__device__ __forceinline__ float reduce8_transposed(const float (&g)[NV], int lane) {
    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
    float r4[4], r2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = u16 ? g[i] : g[i + 4], keep = u16 ? g[i + 4] : g[i];
        r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = u8 ? r4[i] : r4[i + 2], keep = u8 ? r4[i + 2] : r4[i];
        r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = u4 ? r2[0] : r2[1], keep = u4 ? r2[1] : r2[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

__global__ void __launch_bounds__(RENDER_THREADS, 3)
render_backward_kernel(View v, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                       const float4* __restrict__ rec2, const uint32_t* __restrict__ starts,
                       const uint32_t* __restrict__ points, const float* __restrict__ final_T,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dout,
                       float* __restrict__ scratch) {
    __shared__ __align__(16) unsigned char srec[RENDER_THREADS * REC_BYTES];
    __shared__ uint32_t sid[RENDER_THREADS];
    const uint32_t sbase = smem_addr(srec);
    __shared__ uint32_t block_last_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int bx0 = blockIdx.x * TILE + (warp & 1) * 8, by0 = blockIdx.y * TILE + (warp >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = px < v.W && py < v.H;
    const float pxf = (float)px, pyf = (float)py;
    const float wcx = (float)bx0 + 3.5f, wcy = (float)by0 + 1.5f;
    const uint32_t start = starts[tile];

    float Tfin = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
    uint32_t last = 0;
    if (inside) {
        const size_t pix = (size_t)py * v.W + px, hw = (size_t)v.H * v.W;
        Tfin = final_T[pix];
        last = n_contrib[pix];
        d0 = dL_dout[pix];
        d1 = dL_dout[hw + pix];
        d2 = dL_dout[2 * hw + pix];
    }
    // pixels with a zero upstream gradient contribute nothing (crop training leaves most tiles empty)
    if (d0 == 0.f && d1 == 0.f && d2 == 0.f) last = 0;
    if (tid == 0) block_last_s = 0;
    __syncthreads();
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0 && warp_last > 0) atomicMax(&block_last_s, warp_last);
    __syncthreads();
    const uint32_t block_last = block_last_s;
    if (block_last == 0) return;

    const float bg_dot = v.bg[0] * d0 + v.bg[1] * d1 + v.bg[2] * d2;
    const float neg_half_w = -0.5f * (float)v.W, neg_half_h = -0.5f * (float)v.H;
    const float tb = -Tfin * bg_dot;
    float T = Tfin;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
    // destination of this lane after the butterfly: lane 4k owns slot k (k < 8), lane 1 owns slot 8
    const int my_slot = (lane == 1) ? 8 : (lane >> 2);
    const bool commits = ((lane & 3) == 0) || lane == 1;

    const int nb = (int)((block_last + RENDER_THREADS - 1) / RENDER_THREADS);
    for (int bi = nb - 1; bi >= 0; --bi) {
        const uint32_t boff = (uint32_t)bi * RENDER_THREADS;
        const uint32_t cnt = min((uint32_t)RENDER_THREADS, block_last - boff);
        __syncthreads();  // every warp is done with the previous batch before the refill
        if (tid < cnt) {
            const uint32_t id = points[start + boff + tid];
            sid[tid] = id;
            const uint32_t dst = sbase + tid * REC_BYTES;
            sts128(dst, rec0[id]);
            sts128(dst + 16, rec1[id]);
            sts128(dst + 32, rec2[id]);
        }
        __syncthreads();
        if (warp_last <= boff) continue;

        for (int r = (int)((cnt - 1) & ~31u); r >= 0; r -= 32) {
            const uint32_t j = (uint32_t)r + lane;
            bool hit = false;
            if (j < cnt) {
                const float4 a = lds128(sbase + j * REC_BYTES);
                hit = (fabsf(a.x - wcx) <= a.z + 3.5f) && (fabsf(a.y - wcy) <= a.w + 1.5f);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int b = 31 - __clz(mask);
                mask &= ~(1u << b);
                const uint32_t jj = (uint32_t)r + b;
                const uint32_t pos = boff + jj;  // 0-based list position; contributor id is pos+1
                float g[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) g[k] = 0.f;
                bool act = false;
                if (pos < last) {
                    const uint32_t src = sbase + jj * REC_BYTES;
                    const float2 xy = lds64(src);
                    const float4 c = lds128(src + 16);
                    const float dx = xy.x - pxf, dy = xy.y - pyf;
                    const float dxx = dx * dx, dyy = dy * dy, dxy = dx * dy;
                    const float power = -0.5f * (c.x * dxx + c.z * dyy) - c.y * dxy;
                    if (power <= 0.0f) {
                        const float G = ex2_approx(power * LOG2E);
                        const float alpha = fminf(ALPHA_MAX, c.w * G);
                        if (alpha >= ALPHA_MIN) {
                            act = true;
                            const float4 col = lds128(src + 32);
                            const float inv_om = rcp_approx(1.0f - alpha);
                            T *= inv_om;
                            const float w = alpha * T;
                            const float keep = 1.0f - last_alpha;
                            acc0 = fmaf(last_alpha, lc0, keep * acc0);
                            acc1 = fmaf(last_alpha, lc1, keep * acc1);
                            acc2 = fmaf(last_alpha, lc2, keep * acc2);
                            lc0 = col.x, lc1 = col.y, lc2 = col.z;
                            last_alpha = alpha;
                            float dL_dalpha = (col.x - acc0) * d0;
                            dL_dalpha = fmaf(col.y - acc1, d1, dL_dalpha);
                            dL_dalpha = fmaf(col.z - acc2, d2, dL_dalpha);
                            dL_dalpha = fmaf(dL_dalpha, T, tb * inv_om);  // + (-T_final / (1 - alpha)) * bg . dL/dpix
                            const float q = G * dL_dalpha;   // dL/dopacity contribution
                            const float t = c.w * q;          // dL/dG * G
                            const float h = -0.5f * t;
                            g[G_MX] = (t * neg_half_w) * fmaf(c.x, dx, c.y * dy);
                            g[G_MY] = (t * neg_half_h) * fmaf(c.z, dy, c.y * dx);
                            g[G_CA] = h * dxx;
                            g[G_CB] = h * dxy;
                            g[G_CC] = h * dyy;
                            g[G_OP] = q;
                            g[G_R] = w * d0;
                            g[G_G] = w * d1;
                            g[G_B] = w * d2;
                        }
                    }
                }
                if (__any_sync(0xffffffffu, act)) {
                    const float r8 = reduce8_transposed(g, lane);  // lane l: total of value l >> 2
                    const float r9 = warp_sum(g[8]);
                    if (commits) red_add(scratch + (size_t)sid[jj] * GRAD_STRIDE + my_slot, lane == 1 ? r9 : r8);
                }
            }
        }
    }
}

void launch_render_backward(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, const float* dL_dout, float* scratch,
                            cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    render_backward_kernel<<<grid, RENDER_THREADS, 0, s>>>(v, g.rec0, g.rec1, g.rec2, im.starts, b.points, im.final_T,
                                                           im.n_contrib, dL_dout, scratch);
}

}  // namespace ggrt
