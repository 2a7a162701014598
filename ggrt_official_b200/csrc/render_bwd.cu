// Per-tile back-to-front gradient replay (A.4).  Replaces upstream renderCUDA (bwd),
// SURVEY.md 8a row a12 -- the dominant cost of fwd+bwd upstream because every pixel
// issues ~10 global float atomics per contributing Gaussian.
//
// Same tile / warp / batch structure and the same per-warp bounding-box cull as the
// forward.  Per (warp, Gaussian) the 9 partial gradients are reduced over the 32 pixels
// with a transposed shuffle butterfly (14 shuffles for 9 values), added into a per-batch
// shared-memory accumulator, and the CTA flushes one set of global reductions per
// (tile, Gaussian).
#include "common.cuh"

namespace ggrt {

constexpr int RENDER_THREADS = 256;
constexpr int NV = 9;  // gradient values per Gaussian (GradSlot)

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

__global__ void __launch_bounds__(RENDER_THREADS)
render_backward_kernel(View v, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                       const float4* __restrict__ rec2, const uint32_t* __restrict__ starts,
                       const uint32_t* __restrict__ points, const float* __restrict__ final_T,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dout,
                       float* __restrict__ scratch) {
    __shared__ float4 s0[RENDER_THREADS], s1[RENDER_THREADS], s2[RENDER_THREADS];
    __shared__ uint32_t sid[RENDER_THREADS];
    __shared__ float sg[RENDER_THREADS * NV];
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
    const float half_w = 0.5f * (float)v.W, half_h = 0.5f * (float)v.H;
    float T = Tfin;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;

    const int nb = (int)((block_last + RENDER_THREADS - 1) / RENDER_THREADS);
    for (int bi = nb - 1; bi >= 0; --bi) {
        const uint32_t boff = (uint32_t)bi * RENDER_THREADS;
        const uint32_t cnt = min((uint32_t)RENDER_THREADS, block_last - boff);
        __syncthreads();  // previous batch fully flushed before the refill
        if (tid < cnt) {
            const uint32_t id = points[start + boff + tid];
            sid[tid] = id;
            s0[tid] = rec0[id];
            s1[tid] = rec1[id];
            s2[tid] = rec2[id];
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) sg[tid * NV + k] = 0.f;
        __syncthreads();

        if (warp_last > boff) {
            for (int r = (int)((cnt - 1) & ~31u); r >= 0; r -= 32) {
                const uint32_t j = (uint32_t)r + lane;
                bool hit = false;
                if (j < cnt) {
                    const float4 a = s0[j];
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
                        const float4 a = s0[jj], c = s1[jj];
                        const float dx = a.x - pxf, dy = a.y - pyf;
                        const float power = -0.5f * (c.x * dx * dx + c.z * dy * dy) - c.y * dx * dy;
                        if (power <= 0.0f) {
                            const float G = __expf(power);
                            const float alpha = fminf(ALPHA_MAX, c.w * G);
                            if (alpha >= ALPHA_MIN) {
                                act = true;
                                const float4 col = s2[jj];
                                const float om = 1.0f - alpha;
                                const float inv_om = __fdividef(1.0f, om);
                                T = T * inv_om;
                                const float w = alpha * T;
                                acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0;
                                acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1;
                                acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2;
                                lc0 = col.x, lc1 = col.y, lc2 = col.z;
                                float dL_dalpha = (col.x - acc0) * d0 + (col.y - acc1) * d1 + (col.z - acc2) * d2;
                                dL_dalpha *= T;
                                last_alpha = alpha;
                                dL_dalpha += (-Tfin * inv_om) * bg_dot;
                                const float dL_dG = c.w * dL_dalpha;
                                const float gdx = G * dx, gdy = G * dy;
                                const float dG_ddelx = -gdx * c.x - gdy * c.y;
                                const float dG_ddely = -gdy * c.z - gdx * c.y;
                                g[G_MX] = dL_dG * dG_ddelx * half_w;
                                g[G_MY] = dL_dG * dG_ddely * half_h;
                                g[G_CA] = -0.5f * gdx * dx * dL_dG;
                                g[G_CB] = -0.5f * gdx * dy * dL_dG;
                                g[G_CC] = -0.5f * gdy * dy * dL_dG;
                                g[G_OP] = G * dL_dalpha;
                                g[G_R] = w * d0;
                                g[G_G] = w * d1;
                                g[G_B] = w * d2;
                            }
                        }
                    }
                    if (__any_sync(0xffffffffu, act)) {
                        const float r = reduce8_transposed(g, lane);  // lane l: total of value l >> 2
                        const float r8 = warp_sum(g[8]);
                        if ((lane & 3) == 0) atomicAdd(&sg[jj * NV + (lane >> 2)], r);
                        if (lane == 1) atomicAdd(&sg[jj * NV + 8], r8);
                    }
                }
            }
        }
        __syncthreads();
        if (tid < cnt) {
            float* dst = scratch + (size_t)sid[tid] * GRAD_STRIDE;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const float x = sg[tid * NV + k];
                if (x != 0.f) atomicAdd(dst + k, x);
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
