// Micro-benchmark of L2 atomic throughput against counter layout on sm_100a (measurement aid for the binning kernels,
// DESIGN.md section 8; not product code).  800 K atomicAdd on 48 K counters (the C2 tile binning: 3024 tiles x 16
// sub-counters, ~17 increments each), issued like emit_kernel does (300 K threads, 1-4 atomics in flight per thread),
// for counter strides of 4 .. 128 bytes, returning (ATOM) and non-returning (RED).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_variants/ubench_atomics tools/ubench/atomics.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <bool RETURNING>
__global__ void kern(uint32_t* counters, int stride_words, int tiles, int P, uint32_t* sink) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t h = hash(i);
    const int n = (h & 3) == 0 ? 1 : (h & 3) == 1 ? 2 : 4;           // ~2.7 tiles per Gaussian
    const int gx = 63, tx = (h >> 4) % (gx - 1), ty = (h >> 12) % (tiles / gx - 1);
    const int sub = i & 15;
    uint32_t acc = 0;
    uint32_t slot[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        slot[u] = 0;
        if (u < n) {
            const int t = (ty + (u >> 1)) * gx + tx + (u & 1);
            uint32_t* a = counters + (size_t)(t * 16 + sub) * stride_words;
            if (RETURNING) slot[u] = atomicAdd(a, 1u);
            else atomicAdd(a, 1u);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += slot[u];
    if (RETURNING && acc == 0xffffffffu) sink[0] = acc;
}

int main() {
    const int tiles = 3024, P = 300000;
    uint32_t *c, *sink;
    cudaMalloc(&c, (size_t)tiles * 16 * 128);
    cudaMalloc(&sink, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    for (int returning = 0; returning < 2; ++returning)
        for (int stride = 4; stride <= 128; stride *= 2) {
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaMemset(c, 0, (size_t)tiles * 16 * stride);
                cudaDeviceSynchronize();
                cudaEventRecord(a);
                if (returning) kern<true><<<(P + 255) / 256, 256>>>(c, stride / 4, tiles, P, sink);
                else kern<false><<<(P + 255) / 256, 256>>>(c, stride / 4, tiles, P, sink);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("{\"test\": \"%s\", \"counter_stride_bytes\": %d, \"us\": %.2f}\n", returning ? "atom_returning" : "red", stride, best * 1e3f);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
