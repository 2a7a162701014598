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

}  // namespace ggrt
