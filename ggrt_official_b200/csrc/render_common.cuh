// Helpers shared by the two render kernels: approximate SFU ops, explicit shared-memory
// addressing (one 48-byte record per staged Gaussian) and the per-warp pixel-block layout.
#pragma once
#include "common.cuh"

namespace ggrt {

// A tile has 8 warp pixel blocks (8x4 px each).  Each render CTA handles *_WARPS of them; splitting a tile over
// several smaller CTAs shortens the work unit and with it the tail of the last wave.
#ifndef GGRT_FWD_WARPS
#define GGRT_FWD_WARPS 8
#endif
#ifndef GGRT_BWD_WARPS
#define GGRT_BWD_WARPS 8
#endif
constexpr int FWD_WARPS = GGRT_FWD_WARPS, BWD_WARPS = GGRT_BWD_WARPS;
constexpr int FWD_THREADS = 32 * FWD_WARPS, BWD_THREADS = 32 * BWD_WARPS;
constexpr int FWD_BATCH = 256;
constexpr int REC_BYTES = 48;  // {x, y, tau', - | A, B, C, opacity | r, g, b, depth}
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// cp.async (SASS: LDGSTS): 16-byte global -> shared copies without a register round trip, tracked in commit groups
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Exact (conservative) test: does the ellipse {q(p - g) <= tau'} reach the pixel rectangle
// [x0, x0+w] x [y0, y0+h]?  q is convex with its minimum at g, so its minimum over the rectangle is 0 when g is
// inside and otherwise lies on an edge facing g: minimise the 1-D quadratic along the (at most two) facing edges.
// Run by one lane per candidate Gaussian in the cull phase, so its ~25 instructions are amortised over the whole
// pixel block.
__device__ __forceinline__ bool ellipse_hits_rect(float gx, float gy, float tau_c, float A, float B, float C,
                                                  float x0, float y0, float w, float h) {
    const float lox = x0 - gx, hix = lox + w, loy = y0 - gy, hiy = loy + h;  // rectangle relative to the centre
    const float dxe = fminf(fmaxf(0.f, lox), hix), dye = fminf(fmaxf(0.f, loy), hiy);  // nearest point, per axis
    float qmin = 0.f;
    if (dxe != 0.f || dye != 0.f) {
        qmin = 3.0e38f;
        if (dye != 0.f) {  // horizontal edge y = dye: free dx in [lox, hix]
            const float dx = fminf(fmaxf(-B * dye * rcp_approx(A), lox), hix);
            qmin = A * dx * dx + 2.0f * B * dx * dye + C * dye * dye;
        }
        if (dxe != 0.f) {  // vertical edge x = dxe: free dy in [loy, hiy]
            const float dy = fminf(fmaxf(-B * dxe * rcp_approx(C), loy), hiy);
            qmin = fminf(qmin, A * dxe * dxe + 2.0f * B * dxe * dy + C * dy * dy);
        }
    }
    return qmin <= tau_c;
}


// The same test for all 8 warp pixel blocks of a tile at once (block w: columns 8 (w & 1) .. + 7, rows 4 (w >> 1) .. + 3
// from the tile origin), run by ONE thread per staged Gaussian: bit w of the result is set if the ellipse reaches block
// w.  The per-column and per-row terms are shared, so a block costs ~12 instructions.  min(q on the facing vertical
// edge, q on the facing horizontal edge) is the minimum over the rectangle in every case: with the centre inside the
// rectangle's x range the "vertical edge" degenerates to the line x = 0 through the centre, whose minimum over the
// row range is not above the other candidate (and 0 when the centre is inside).
__device__ __forceinline__ uint32_t block_mask8(float gx, float gy, float tau_c, float A, float B, float C, float tx0,
                                                float ty0) {
    const float iA = rcp_approx(A), iC = rcp_approx(C), B2 = 2.0f * B;
    float lox[2], hix[2], dxe[2], qxe[2], dy0[2], bxe[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        lox[c] = tx0 + 8.0f * c - gx, hix[c] = lox[c] + 7.0f;
        dxe[c] = fminf(fmaxf(0.f, lox[c]), hix[c]);
        qxe[c] = A * dxe[c] * dxe[c], bxe[c] = B2 * dxe[c], dy0[c] = -B * dxe[c] * iC;
    }
    uint32_t mask = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float loy = ty0 + 4.0f * r - gy, hiy = loy + 3.0f;
        const float dye = fminf(fmaxf(0.f, loy), hiy);
        const float qye = C * dye * dye, bye = B2 * dye, dx0 = -B * dye * iA;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float dy = fminf(fmaxf(dy0[c], loy), hiy), dx = fminf(fmaxf(dx0, lox[c]), hix[c]);
            const float qv = fmaf(fmaf(C, dy, bxe[c]), dy, qxe[c]);  // A dxe^2 + 2 B dxe dy + C dy^2
            const float qh = fmaf(fmaf(A, dx, bye), dx, qye);        // C dye^2 + 2 B dye dx + A dx^2
            if (fminf(qv, qh) <= tau_c) mask |= 1u << (2 * r + c);
        }
    }
    return mask;
}

}  // namespace ggrt
