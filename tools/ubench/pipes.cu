// Micro-benchmark of the SM issue pipes on sm_100a (measurement aid for DESIGN.md section 8; not product code).
// One CTA per SM, W warps per CTA, each warp runs a long unrolled sequence of independent instructions of one
// kind (or a fixed mix) and reports cycles per warp-instruction per SM sub-partition (4 per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_variants/ubench_pipes tools/ubench/pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 256
#define UNROLL 8   // independent chains per thread

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 d; asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(v)); return d; }

template <int MODE>
__global__ void kern(float* out, long long* cyc, float a, float b) {
    float x[UNROLL];
    u64 y[UNROLL];
    int z[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) x[i] = threadIdx.x * 0.001f + i, y[i] = pk(x[i], x[i] + 1.f), z[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);                                      // FFMA (2 regs + ... )
            if (MODE == 1) x[i] = fmaf(x[i], x[(i + 1) % UNROLL], x[(i + 3) % UNROLL]);  // FFMA, 3 distinct regs
            if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(y[(i + 1) % UNROLL]), "l"(y[(i + 3) % UNROLL]));
            if (MODE == 3) {                                                             // FFMA + IADD pairs
                x[i] = fmaf(x[i], a, b);
                z[i] = z[i] * 3 + it;
            }
            if (MODE == 4) {                                                             // FFMA2 + LOP/IADD
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(y[(i + 1) % UNROLL]), "l"(y[(i + 3) % UNROLL]));
                z[i] = (z[i] ^ it) + i;
            }
            if (MODE == 5) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1, 8);               // SHFL
            if (MODE == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));      // MUFU
            if (MODE == 7) {                                                             // FFMA + SHFL 4:1
                x[i] = fmaf(x[i], a, b);
                if ((i & 3) == 0) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1, 8);
            }
            if (MODE == 8) x[i] = fminf(x[i] + a, b);                                    // FADD + FMNMX (fma + alu pipe)
            if (MODE == 10) {                                                            // HMMA.1688.F32.TF32, 8 independent accumulators
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(x[i]), "+f"(x[(i + 1) % UNROLL]), "+f"(x[(i + 2) % UNROLL]), "+f"(x[(i + 3) % UNROLL])
                             : "r"(z[i]), "r"(z[(i + 1) % UNROLL]), "r"(z[(i + 2) % UNROLL]), "r"(z[(i + 3) % UNROLL]), "r"(z[(i + 4) % UNROLL]), "r"(z[(i + 5) % UNROLL]));
            }
            if (MODE == 11) {                                                            // 1 HMMA per 16 FFMA2 (the K7 mix)
                if (i == 0) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(x[0]), "+f"(x[1]), "+f"(x[2]), "+f"(x[3])
                             : "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(y[(i + 1) % UNROLL]), "l"(y[(i + 3) % UNROLL]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(y[(i + 2) % UNROLL]), "l"(y[(i + 5) % UNROLL]));
            }
            if (MODE == 9) {                                                             // 2 FFMA vs. the same on FFMA2: compare 0 with 2
                x[i] = fmaf(x[i], a, b);
                x[i] = fmaf(x[i], b, a);
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s += x[i] + upk(y[i]).x + upk(y[i]).y + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, float* out, long long* cyc) {
    for (int warps : {4, 8, 16, 32}) {
        kern<MODE><<<148, warps * 32>>>(out, cyc, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        // warp-instructions issued per SM sub-partition: (warps / 4) * ITERS * UNROLL * instr_per_iter
        double per = avg / ((warps / 4.0) * ITERS * UNROLL * instr_per_iter);
        printf("{\"test\": \"%s\", \"warps_per_sm\": %d, \"cycles_per_warp_instr_per_smsp\": %.3f}\n", name, warps, per);
    }
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    run<0>("ffma_imm_like(x*a+b)", 1, out, cyc);
    run<1>("ffma_3reg", 1, out, cyc);
    run<2>("ffma2_3reg", 1, out, cyc);
    run<3>("ffma+imad(per pair of 2)", 2, out, cyc);
    run<4>("ffma2+lop+iadd(per 3)", 3, out, cyc);
    run<5>("shfl_up", 1, out, cyc);
    run<6>("mufu_ex2", 1, out, cyc);
    run<7>("ffma+shfl 4:1 (per 1.25)", 1, out, cyc);
    run<8>("fadd+fmnmx(per 2)", 2, out, cyc);
    run<9>("2xffma dependent(per 2)", 2, out, cyc);
    run<10>("hmma_1688_tf32 (4 overlapping accumulator quads)", 1, out, cyc);
    run<11>("1 hmma + 16 ffma2 (per 17/8 per unroll slot)", 2, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
