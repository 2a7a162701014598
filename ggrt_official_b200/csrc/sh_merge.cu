// Exchange step of the view-sharded multi-GPU path: every GPU renders another target view of the
// same Gaussians and the per-view Gaussian gradients must be summed.  88 % of that gradient is dL/dsh
// (12K = 300 B of the 340 B per Gaussian at SH degree 4) -- but dL/dsh of ONE view is a rank-1 object, the outer
// product basis(dir_v) x dL/drgb_v, and dir_v only depends on the Gaussian's position and that view's camera
// centre, which every rank knows.  So the ranks exchange the 12-byte colour gradients instead of the 300-byte SH
// gradients and each rebuilds
//     dL/dsh[i] = sum_v basis(normalize(s * mean_i - campos_v)) (x) dL/drgb_v[i]
// locally.  NVLink traffic per GPU drops from 2 x 340 B (all-reduce) to 12 B x (views - 1) + 2 x 40 B per
// Gaussian; the result is the same sum in another association order (float32).
//
// sh_gradient_merge_kernel reads the per-view colour gradients through plain device pointers, which may be
// PEER pointers into the other GPUs' symmetric buffers: the gather over NVLink then happens inside this kernel
// (all loads of a Gaussian are issued before any is used) and overlaps the HBM write of dL/dsh, slab by slab.
// The slab (128 Gaussians x 12K B, contiguous in HBM) is assembled in shared memory, one odd-stride row per
// thread, and written with a TMA bulk store while the next slab is computed.
//
// nvls_allreduce_kernel sums the small remaining arena [P, 3 + 6 + 1] in place over an NVLS multicast mapping:
// rank r reduces slice r inside the switch (multimem.ld_reduce) and broadcasts it (multimem.st) -- every byte
// crosses NVLink once in each direction.
#include "common.cuh"
#include "tma.cuh"

namespace ggrt {

// Tuned on B200 at C2 (round 2, profiles/r2_merge_variants.json): the kernel is bound by the latency of its
// per-view basis evaluation, not by HBM, so warps per SM matter more than a second ring stage -- one stage, 4 views
// prefetched, <= 128 registers (4 CTAs / SM): 47.2 -> 35.0 us at 8 views, 24.8 -> 23.5 us at 1 view.
#ifndef GGRT_MERGE_STAGES
#define GGRT_MERGE_STAGES 1
#endif
#ifndef GGRT_MERGE_GROUP
#define GGRT_MERGE_GROUP 4
#endif
#ifndef GGRT_MERGE_MINBLOCKS
#define GGRT_MERGE_MINBLOCKS 4
#endif
#ifndef GGRT_MERGE_THREADS
#define GGRT_MERGE_THREADS 128
#endif
constexpr int MERGE_THREADS = GGRT_MERGE_THREADS;
constexpr int MERGE_STAGES = GGRT_MERGE_STAGES;
constexpr int MERGE_GROUP = GGRT_MERGE_GROUP;  // views whose loads are in flight together

struct MergeViews {
    const float* drgb[GGRT_RASTER_MAX_MERGE_VIEWS];    // [P,3] each (device or peer memory)
    const float* campos[GGRT_RASTER_MAX_MERGE_VIEWS];  // [3] each
    int n;
};

// system-scope relaxed load: the data may live on another GPU and was published by a cross-GPU barrier
__device__ __forceinline__ float ld_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

template <int DEG, bool CMAJOR>
__global__ void __launch_bounds__(MERGE_THREADS, GGRT_MERGE_MINBLOCKS)
sh_gradient_merge_kernel(int P, float scale, const float* __restrict__ means, MergeViews mv, float* __restrict__ dsh,
                         int num_slabs, MergeSignal sig) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    constexpr int ROW = 3 * K;
    constexpr int KS = CMAJOR ? 1 : 3, CS = CMAJOR ? K : 1;  // element (k, c) of a row lives at k*KS + c*CS
    extern __shared__ __align__(128) float merge_ring[];
    __shared__ float scam[GGRT_RASTER_MAX_MERGE_VIEWS][3];
    __shared__ long long poff_s;
    // Signalled exchange: the slots are filled by the peers' backward kernels (pushed over NVLink); wait until all
    // `world` ranks have signalled the step this GPU's own backward just completed (*epoch), then read that half.
    if (threadIdx.x == 0) {
        long long poff = 0;
        if (sig.epoch != nullptr) {
            const uint32_t e = *reinterpret_cast<const volatile uint32_t*>(sig.epoch);
            wait_reached(sig.arrive, e * (uint32_t)sig.world);
            poff = (long long)((e - 1u) & 1u) * sig.parity_stride;
        }
        poff_s = poff;
    }
    __syncthreads();
    const long long poff = poff_s;
    if (threadIdx.x < 3 * mv.n)
        scam[threadIdx.x / 3][threadIdx.x % 3] = ld_sys(mv.campos[threadIdx.x / 3] + poff + threadIdx.x % 3);
    __syncthreads();
    const bool aligned = (reinterpret_cast<uintptr_t>(dsh) & 15) == 0;

    // The inputs of the NEXT slab (mean + the first MERGE_GROUP views' colour gradients, possibly remote: NVLink
    // latency is microseconds) are loaded into registers while the current slab is evaluated and stored.
    float n_m[3] = {0.f, 0.f, 0.f};
    float n_g[MERGE_GROUP][3];
    auto prefetch = [&](int sl) {
        const int i = sl * MERGE_THREADS + threadIdx.x;
        const bool ok = sl < num_slabs && i < P;
#pragma unroll
        for (int u = 0; u < MERGE_GROUP; ++u) {
            n_g[u][0] = n_g[u][1] = n_g[u][2] = 0.f;
            if (ok && u < mv.n) {
                const float* src = mv.drgb[u] + poff + 3 * (size_t)i;
                n_g[u][0] = ld_sys(src), n_g[u][1] = ld_sys(src + 1), n_g[u][2] = ld_sys(src + 2);
            }
        }
        if (ok) n_m[0] = means[3 * (size_t)i], n_m[1] = means[3 * (size_t)i + 1], n_m[2] = means[3 * (size_t)i + 2];
    };
    prefetch(blockIdx.x);

    int it = 0;
    for (int sl = blockIdx.x; sl < num_slabs; sl += gridDim.x, ++it) {
        float* slab = merge_ring + (it % MERGE_STAGES) * (MERGE_THREADS * ROW);
        const int base = sl * MERGE_THREADS;
        const int cnt = min(MERGE_THREADS, P - base);
        const int i = base + threadIdx.x;
        const bool valid = threadIdx.x < cnt;
        const float mx = fmul(n_m[0], scale), my = fmul(n_m[1], scale), mz = fmul(n_m[2], scale);
        float g[MERGE_GROUP][3];
#pragma unroll
        for (int u = 0; u < MERGE_GROUP; ++u) g[u][0] = n_g[u][0], g[u][1] = n_g[u][1], g[u][2] = n_g[u][2];
        prefetch(sl + gridDim.x);
        if (threadIdx.x == 0) bulk_wait_read<MERGE_STAGES - 1>();  // the store that last used this stage has drained
        __syncthreads();

        float acc[ROW];
#pragma unroll
        for (int k = 0; k < ROW; ++k) acc[k] = 0.f;
        for (int v0 = 0; v0 < mv.n; v0 += MERGE_GROUP) {
            if (v0 > 0) {  // more than MERGE_GROUP views: the later groups are loaded in place
#pragma unroll
                for (int u = 0; u < MERGE_GROUP; ++u) {
                    g[u][0] = g[u][1] = g[u][2] = 0.f;
                    if (valid && v0 + u < mv.n) {
                        const float* src = mv.drgb[v0 + u] + poff + 3 * (size_t)i;
                        g[u][0] = ld_sys(src), g[u][1] = ld_sys(src + 1), g[u][2] = ld_sys(src + 2);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < MERGE_GROUP; ++u) {
                if (v0 + u >= mv.n) break;
                const float dR = g[u][0], dG = g[u][1], dB = g[u][2];
                if (dR == 0.f && dG == 0.f && dB == 0.f) continue;  // culled in this view (or no gradient)
                const float vx = mx - scam[v0 + u][0], vy = my - scam[v0 + u][1], vz = mz - scam[v0 + u][2];
                const float inv = rsqrtf(vx * vx + vy * vy + vz * vz);
                float b[K];
                sh_basis(DEG, vx * inv, vy * inv, vz * inv, b);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    acc[k * KS] = fmaf(b[k], dR, acc[k * KS]);
                    acc[k * KS + CS] = fmaf(b[k], dG, acc[k * KS + CS]);
                    acc[k * KS + 2 * CS] = fmaf(b[k], dB, acc[k * KS + 2 * CS]);
                }
            }
        }
        if (valid) {
            float* my_row = slab + threadIdx.x * ROW;
#pragma unroll
            for (int k = 0; k < ROW; ++k) my_row[k] = acc[k];
        }
        float* dst = dsh + (size_t)base * ROW;
        const int nfl = cnt * ROW;
        if (aligned && ((nfl * 4) & 15) == 0) {
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_s2g(dst, smem_u32(slab), (uint32_t)nfl * 4u);
                bulk_commit();
            }
        } else {  // ragged last slab / unaligned output
            __syncthreads();
            for (int k = threadIdx.x; k < nfl; k += MERGE_THREADS) dst[k] = slab[k];
        }
    }
    if (threadIdx.x == 0) bulk_wait0();
}

template <int DEG>
static void launch_merge_deg(int P, float scale, bool cmajor, const float* means, const MergeViews& mv, float* dsh,
                             const MergeSignal& sig, cudaStream_t s) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    const size_t smem = (size_t)MERGE_STAGES * MERGE_THREADS * 3 * K * sizeof(float);
    const int num_slabs = (P + MERGE_THREADS - 1) / MERGE_THREADS;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = max(1, min(GGRT_MERGE_MINBLOCKS, (int)((220 * 1024) / (smem + 1024))));
    const int grid = min(num_slabs, per_sm * sms);
    if (cmajor && K > 1) {
        if (smem > 32 * 1024)
            cudaFuncSetAttribute(sh_gradient_merge_kernel<DEG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sh_gradient_merge_kernel<DEG, true><<<grid, MERGE_THREADS, smem, s>>>(P, scale, means, mv, dsh, num_slabs, sig);
    } else {
        if (smem > 32 * 1024)
            cudaFuncSetAttribute(sh_gradient_merge_kernel<DEG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sh_gradient_merge_kernel<DEG, false><<<grid, MERGE_THREADS, smem, s>>>(P, scale, means, mv, dsh, num_slabs, sig);
    }
}

void launch_sh_gradient_merge(int P, int deg, float scale, bool cmajor, const float* means, int num_views,
                              const float* const* drgb, const float* const* campos, float* dsh, const MergeSignal& sig,
                              cudaStream_t s) {
    if (P == 0) return;
    MergeViews mv;
    mv.n = num_views;
    for (int v = 0; v < GGRT_RASTER_MAX_MERGE_VIEWS; ++v) {
        mv.drgb[v] = v < num_views ? drgb[v] : nullptr;
        mv.campos[v] = v < num_views ? campos[v] : nullptr;
    }
    switch (deg) {
        case 0: launch_merge_deg<0>(P, scale, cmajor, means, mv, dsh, sig, s); break;
        case 1: launch_merge_deg<1>(P, scale, cmajor, means, mv, dsh, sig, s); break;
        case 2: launch_merge_deg<2>(P, scale, cmajor, means, mv, dsh, sig, s); break;
        case 3: launch_merge_deg<3>(P, scale, cmajor, means, mv, dsh, sig, s); break;
        default: launch_merge_deg<4>(P, scale, cmajor, means, mv, dsh, sig, s); break;
    }
}

// ---- in-place sum over ranks through an NVLS multicast mapping (two-shot: reduce my slice, broadcast it) ----
// Signalled form (epoch != NULL): both cross-GPU waits are inside the kernel -- it starts when every rank's inputs
// are complete (arrive_in reached world * *epoch) and its last CTA returns when every rank has broadcast its slice.
__global__ void __launch_bounds__(512)
nvls_allreduce_kernel(float* __restrict__ mc, long long first4, long long n4, const uint32_t* epoch, int world,
                      const uint32_t* arrive_in, uint32_t* arrive_out_mc, const uint32_t* arrive_out,
                      uint32_t* done_counter) {
    uint32_t target = 0;
    if (epoch != nullptr) {
        if (threadIdx.x == 0) {
            target = *reinterpret_cast<const volatile uint32_t*>(epoch) * (uint32_t)world;
            wait_reached(arrive_in, target);
        }
        __syncthreads();
    }
    float4* p = reinterpret_cast<float4*>(mc) + first4;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n4; k += (long long)gridDim.x * blockDim.x) {
        float4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(p + k)
                     : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + k), "f"(v.x), "f"(v.y),
                     "f"(v.z), "f"(v.w)
                     : "memory");
    }
    if (epoch != nullptr && last_cta_done(done_counter)) {
        signal_add(arrive_out_mc, true);    // this rank's slice is in every GPU's buffer
        wait_reached(arrive_out, target);   // ... and so are the slices of all other ranks
    }
}

// Cross-GPU barrier in one single-thread kernel: every rank adds 1 to ALL copies of a counter with one
// multimem.red through the NVLS multicast mapping (release: the pushes of the preceding kernels are ordered before
// it), then spins on its own copy until all `world` increments of this epoch have arrived (acquire).
__global__ void nvls_barrier_kernel(unsigned int* mc_counter, const unsigned int* local_counter, unsigned int target) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_counter), "r"(1u) : "memory");
    unsigned int v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local_counter) : "memory");
    } while ((int)(v - target) < 0);
}

void launch_nvls_barrier(unsigned int* mc_counter, const unsigned int* local_counter, unsigned int target,
                         cudaStream_t s) {
    nvls_barrier_kernel<<<1, 32, 0, s>>>(mc_counter, local_counter, target);
}

void launch_nvls_allreduce(float* multicast, long long count, int rank, int world, const uint32_t* epoch,
                           const uint32_t* arrive_in, uint32_t* arrive_out_mc, const uint32_t* arrive_out,
                           uint32_t* done_counter, cudaStream_t s) {
    const long long n4 = count / 4;  // count is a multiple of 4 (checked by the caller)
    const long long per = (n4 + world - 1) / world;
    const long long first = per * rank < n4 ? per * rank : n4, last = per * (rank + 1) < n4 ? per * (rank + 1) : n4;
    if (last <= first && epoch == nullptr) return;  // (a signalled call must still signal with an empty slice)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = last > first ? (last - first + 511) / 512 : 1;
    const int grid = (int)(want < 2LL * sms ? want : 2LL * sms);
    nvls_allreduce_kernel<<<grid, 512, 0, s>>>(multicast, first, last > first ? last - first : 0, epoch, world, arrive_in,
                                               arrive_out_mc, arrive_out, done_counter);
}

}  // namespace ggrt
