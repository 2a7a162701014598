// Packed FP32 pairs (sm_100a FFMA2 / FMUL2 / FADD2): one issue slot does two float operations.  The render kernels
// are instruction-issue bound, so every pair of independent float operations that can be expressed this way halves
// its issue cost.  A scalar operand is broadcast by packing it twice -- ptxas folds that into the instruction's
// ".F32" operand form, no extra register or move (checked in SASS).
#pragma once
#include <stdint.h>

namespace ggrt {

struct f2 {
    unsigned long long v;
};

__device__ __forceinline__ f2 pk(float a, float b) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ f2 bc(float a) { return pk(a, a); }
__device__ __forceinline__ float lo(f2 a) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
    return x;
}
__device__ __forceinline__ float hi(f2 a) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
    return y;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
    return d;
}
__device__ __forceinline__ f2 shfl_up2(f2 a, int delta, int width) {
    f2 r;
    r.v = __shfl_up_sync(0xffffffffu, a.v, delta, width);
    return r;
}
__device__ __forceinline__ f2 shfl_xor2(f2 a, int mask) {
    f2 r;
    r.v = __shfl_xor_sync(0xffffffffu, a.v, mask);
    return r;
}

}  // namespace ggrt
