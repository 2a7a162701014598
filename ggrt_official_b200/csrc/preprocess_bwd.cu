// Per-Gaussian backward (A.5): conic -> cov2D -> cov3D and mean (through J), NDC mean ->
// mean3D (through the full projection), colour -> SH coefficients and mean3D (through the
// view direction).  Replaces upstream computeCov2DCUDA + preprocessCUDA (bwd), SURVEY.md
// 8a row a13, fused into one pass.  Every output element is written (zeros for culled
// Gaussians) so the caller needs no memset of the 300 B/Gaussian SH gradient.
//
// HBM-bound (about 480 B per Gaussian at SH degree 4, 300 of them the dL/dsh row it writes).  The SH table is NOT
// read: the only use the backward has for the coefficients is d(colour)/d(view direction), a 3x3 matrix per
// Gaussian that the forward's colour kernel stores while the row is on chip (GeomPtrs::jac, 36 B instead of
// 12K B).  Persistent CTAs; the dL/dsh slab of PB_THREADS Gaussians (contiguous in HBM) is assembled in a
// shared-memory ring by the threads (one row each, odd stride) and written with TMA bulk stores
// (cp.async.bulk.global.shared) while the next slab is processed.
//
// Compact mode (shs given, dsh == NULL, dcolors != NULL): dL/dsh of one view is the outer product
// basis(dir) x dL/drgb, so for the view-sharded multi-GPU path only the masked colour gradient [P,3] is
// written (12 B instead of 12K B per Gaussian) and the SH gradient of ALL views is rebuilt after the exchange
// by sh_gradient_merge_kernel (sh_merge.cu).  No shared-memory ring at all in this mode.
#include "common.cuh"
#include "tma.cuh"

namespace ggrt {

#ifndef GGRT_PB_THREADS
#define GGRT_PB_THREADS 32
#endif
#ifndef GGRT_PB_STAGES
#define GGRT_PB_STAGES 2
#endif
constexpr int PB_THREADS = GGRT_PB_THREADS;
#ifndef GGRT_PB_MINBLOCKS
#define GGRT_PB_MINBLOCKS 16  // 128 registers (4 bytes of spill): 16 one-warp CTAs per SM hide the kernel's dependent chains
#endif

constexpr int PB_STAGES = GGRT_PB_STAGES;
#ifndef GGRT_PB_SPLIT_BLOCKS
#define GGRT_PB_SPLIT_BLOCKS 16
#endif

// float4 / scalar stores of the compact colour gradients: plain (local or peer memory) or NVLS multicast
__device__ __forceinline__ void sink_store4(float* dst, float4 v, bool multimem) {
    if (multimem)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y),
                     "f"(v.z), "f"(v.w)
                     : "memory");
    else
        *reinterpret_cast<float4*>(dst) = v;
}
__device__ __forceinline__ void sink_store1(float* dst, float v, bool multimem) {
    if (multimem)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory");
    else
        *dst = v;
}

// (the camera-gradient instantiation carries 27 more accumulators: it keeps the 10-CTA register budget instead of spilling)
template <bool AUX, bool CMAJOR, bool POSE>
__global__ void __launch_bounds__(PB_THREADS, POSE ? (GGRT_PB_MINBLOCKS < 10 ? GGRT_PB_MINBLOCKS : 10) : GGRT_PB_MINBLOCKS)
preprocess_backward_kernel(View v, const float* __restrict__ means, const float* __restrict__ cov3d,
                           const float* __restrict__ shs, const int* __restrict__ radii,
                           const uint8_t* __restrict__ flags, const float* __restrict__ jac, size_t jac_plane,
                           const float* __restrict__ scratch,
                           float* __restrict__ dmeans2D, float* __restrict__ dopacity, float* __restrict__ dmeans3D,
                           float* __restrict__ dcov3D, float* __restrict__ dsh, float* __restrict__ dcolors,
                           float* __restrict__ daux, float* __restrict__ dcam, ColorSinks sinks, int num_slabs) {
    extern __shared__ __align__(128) float slab_ring[];
    __shared__ __align__(16) float sdc[PB_THREADS * 3];  // compact mode: the slab's colour gradients, staged for
                                                          // coalesced 16-byte (possibly remote / multicast) stores
    __shared__ float sV[16], sM[16];
    pdl_enter();
    if (threadIdx.x < 16) sV[threadIdx.x] = v.view[threadIdx.x];
    else if (threadIdx.x < 32) sM[threadIdx.x - 16] = v.proj[threadIdx.x - 16];

    resolve_device_params(v);
    // signalled exchange: this step pushes into half (epoch & 1) of every sink
    const long long poff = sinks.epoch != nullptr
                               ? (long long)(*reinterpret_cast<volatile uint32_t*>(sinks.epoch) & 1u) * sinks.parity_stride
                               : 0ll;
    const int row = v.K * 3;
    const bool compact = shs != nullptr && dsh == nullptr;  // uniform: write dL/drgb instead of dL/dsh
    const bool push = compact && sinks.n > 0;               // (no sink: dL/dsh is written by sh_gradient_kernel)
    const int ks = CMAJOR ? 1 : 3, cs = CMAJOR ? v.K : 1;  // SH element (k, c) at k*ks + c*cs of the row
    const int slab_floats = PB_THREADS * row;
    const bool use_jac = shs != nullptr && v.deg > 0;   // uniform
    const bool tma_ok = shs != nullptr && !compact && (reinterpret_cast<uintptr_t>(dsh) & 15) == 0;
    __syncthreads();  // publishes sV / sM

    // per-Gaussian inputs of the NEXT slab are prefetched into registers (one DRAM latency, overlapped with
    // the evaluation of the current slab) instead of being loaded behind data-dependent branches
    int n_radius = 0;
    uint32_t n_flags = 0;
    float n_mean[3] = {0.f, 0.f, 0.f}, n_cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float4 n_ga = make_float4(0.f, 0.f, 0.f, 0.f), n_gb = n_ga;
    float n_gc = 0.f, n_gx = 0.f;
    float n_jac[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto prefetch = [&](int sl) {
        const int i = sl * PB_THREADS + threadIdx.x;
        n_radius = 0;
        if (sl < num_slabs && i < v.P) {
            n_radius = radii[i];
            n_flags = flags[i];
#pragma unroll
            for (int k = 0; k < 3; ++k) n_mean[k] = means[3 * (size_t)i + k];  // raw: the scene scale is applied
            load_cov6_raw(v, cov3d, i, n_cv);                                  // when the values are consumed
            const float* gs = scratch + (size_t)i * GRAD_STRIDE;
            n_ga = *reinterpret_cast<const float4*>(gs);
            n_gb = *reinterpret_cast<const float4*>(gs + 4);
            n_gc = gs[8];
            if (AUX && (daux || v.aux_mode)) n_gx = gs[G_AUX];
            if (use_jac && n_radius > 0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) n_jac[k] = jac[(size_t)k * jac_plane + i];
            }
        }
    };
    prefetch(blockIdx.x);
    const float cpx = v.campos[0], cpy = v.campos[1], cpz = v.campos[2];

    // opt-in camera gradients (dcam != NULL): per-thread partial sums of dL/dviewmatrix (12 entries: columns
    // 0..2), dL/dprojmatrix (12 entries: columns 0, 1, 3) and dL/dcampos, reduced once at the end of the kernel
    float camV[12], camM[12], camC[3];
#pragma unroll
    for (int k = 0; k < 12; ++k) camV[k] = 0.f, camM[k] = 0.f;
    camC[0] = camC[1] = camC[2] = 0.f;

    int it = 0;
    for (int sl = blockIdx.x; sl < num_slabs; sl += gridDim.x, ++it) {
    const int st = it % PB_STAGES;
    float* slab = slab_ring + st * slab_floats;
    const int base = sl * PB_THREADS;
    const int cnt = min(PB_THREADS, v.P - base);
    const int i = base + threadIdx.x;
    const bool valid = threadIdx.x < cnt;
    const bool vis = valid && n_radius > 0;
    const int nfl = cnt * row;
    const bool slab_tma = tma_ok && ((nfl * 4) & 15) == 0;
    const uint32_t fl = n_flags;
    const float mean_x = fmul(n_mean[0], v.scale), mean_y = fmul(n_mean[1], v.scale), mean_z = fmul(n_mean[2], v.scale);
    float cv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) cv[k] = n_cv[k];
    scale_cov6(v, cv);
    const float4 ga = n_ga, gb = n_gb;
    const float gc = n_gc, gaux = n_gx;
    float jc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) jc[k] = n_jac[k];
    prefetch(sl + gridDim.x);

    // the stage this slab's dL/dsh rows are assembled in must have been drained by the bulk store issued
    // PB_STAGES slabs ago
    if (shs != nullptr && !compact) {
        if (threadIdx.x == 0) bulk_wait_read<PB_STAGES - 1>();
        __syncthreads();
    }

    float dmean[3] = {0.f, 0.f, 0.f};
    float dS[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float g2x = 0.f, g2y = 0.f, gop = 0.f, dR = 0.f, dG = 0.f, dB = 0.f;
    float* my = slab + threadIdx.x * row;

    bool live = false;
    Geo q;
    if (vis) live = geometry(v, sV, sM, mean_x, mean_y, mean_z, cv, q);
    if (live) {
        g2x = ga.x, g2y = ga.y;
        const float gA = ga.z, gBh = ga.w, gC = gb.x;
        gop = gb.y;
        dR = gb.z, dG = gb.w, dB = gc;

        // conic -> cov2D
        const float a = q.a, b = q.b, c = q.c, det = q.det;
        const float k2 = 1.0f / (det * det + 0.0000001f);
        const float dL_da = k2 * (-c * c * gA + 2.0f * b * c * gBh + (det - a * c) * gC);
        const float dL_dc = k2 * (-a * a * gC + 2.0f * a * b * gBh + (det - a * c) * gA);
        const float dL_db = k2 * 2.0f * (b * c * gA - (det + 2.0f * b * b) * gBh + a * b * gC);
        const float(*Tm)[3] = q.Tm;
        // cov2D -> cov3D (xx,xy,xz,yy,yz,zz)
        dS[0] = Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc;
        dS[3] = Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc;
        dS[5] = Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc;
        dS[1] = 2.0f * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db +
                2.0f * Tm[1][0] * Tm[1][1] * dL_dc;
        dS[2] = 2.0f * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db +
                2.0f * Tm[1][0] * Tm[1][2] * dL_dc;
        dS[4] = 2.0f * Tm[0][1] * Tm[0][2] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db +
                2.0f * Tm[1][1] * Tm[1][2] * dL_dc;
        // cov2D -> Tm -> J -> t
        const float S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
        float dTm[2][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float ts0 = Tm[0][0] * S[0][k] + Tm[0][1] * S[1][k] + Tm[0][2] * S[2][k];
            const float ts1 = Tm[1][0] * S[0][k] + Tm[1][1] * S[1][k] + Tm[1][2] * S[2][k];
            dTm[0][k] = 2.0f * dL_da * ts0 + dL_db * ts1;
            dTm[1][k] = dL_db * ts0 + 2.0f * dL_dc * ts1;
        }
        const float dJ00 = dTm[0][0] * sV[0] + dTm[0][1] * sV[4] + dTm[0][2] * sV[8];
        const float dJ02 = dTm[0][0] * sV[2] + dTm[0][1] * sV[6] + dTm[0][2] * sV[10];
        const float dJ11 = dTm[1][0] * sV[1] + dTm[1][1] * sV[5] + dTm[1][2] * sV[9];
        const float dJ12 = dTm[1][0] * sV[2] + dTm[1][1] * sV[6] + dTm[1][2] * sV[10];
        const float limx = 1.3f * v.tanfovx, limy = 1.3f * v.tanfovy;
        const float xmul = (q.txtz < -limx || q.txtz > limx) ? 0.0f : 1.0f;
        const float ymul = (q.tytz < -limy || q.tytz > limy) ? 0.0f : 1.0f;
        const float z1 = 1.0f / q.tz, z2 = z1 * z1, z3 = z2 * z1;
        const float dtx = xmul * -v.fx * z2 * dJ02;
        const float dty = ymul * -v.fy * z2 * dJ12;
        float dtz = -v.fx * z2 * dJ00 - v.fy * z2 * dJ11 + (2.0f * v.fx * q.cx) * z3 * dJ02 +
                    (2.0f * v.fy * q.cy) * z3 * dJ12;
        // GGRt's depth channel is a function of the view depth: its gradient enters through t.z (the final scene-scale
        // chain rule below turns d/d(scaled mean) into d/d(mean))
        if (AUX && v.aux_mode == 1 && ggrt_depth_channel(q.tz, v.scale) > 0.0f) dtz += gaux * GGRT_SH_C0 / v.scale;
        // NDC mean -> mean3D through the full projection
        const float mw = q.pw;
        const float mul1 = q.hx * mw * mw, mul2 = q.hy * mw * mw;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dmean[k] = sV[4 * k + 0] * dtx + sV[4 * k + 1] * dty + sV[4 * k + 2] * dtz +
                       (sM[4 * k + 0] * mw - sM[4 * k + 3] * mul1) * g2x +
                       (sM[4 * k + 1] * mw - sM[4 * k + 3] * mul2) * g2y;
        }
        if (POSE) {
            // t = [p,1].V[:, :3] and hom = [p,1].M  =>  dL/dV[i][j] += p_i dL/dt_j,  dL/dM[i][j] += p_i dL/dhom_j;
            // Tm = J.Rw with Rw[m][k] = V[k][m]    =>  dL/dV[k][m] += sum_r dTm[r][k] J[r][m]
            const float ph[4] = {mean_x, mean_y, mean_z, 1.0f};
            const float dt[3] = {dtx, dty, dtz};
            const float J00 = v.fx * z1, J02 = -v.fx * q.cx * z2, J11 = v.fy * z1, J12 = -v.fy * q.cy * z2;
            const float dhx = g2x * mw, dhy = g2y * mw, dhw = -(q.hx * g2x + q.hy * g2y) * mw * mw;
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
#pragma unroll
                for (int j = 0; j < 3; ++j) camV[3 * i2 + j] = fmaf(ph[i2], dt[j], camV[3 * i2 + j]);
                camM[3 * i2 + 0] = fmaf(ph[i2], dhx, camM[3 * i2 + 0]);
                camM[3 * i2 + 1] = fmaf(ph[i2], dhy, camM[3 * i2 + 1]);
                camM[3 * i2 + 2] = fmaf(ph[i2], dhw, camM[3 * i2 + 2]);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                camV[3 * k + 0] = fmaf(dTm[0][k], J00, camV[3 * k + 0]);
                camV[3 * k + 1] = fmaf(dTm[1][k], J11, camV[3 * k + 1]);
                camV[3 * k + 2] += dTm[0][k] * J02 + dTm[1][k] * J12;
            }
        }
    }

    if (shs != nullptr) {
        if (live) {
            if (fl & 1) dR = 0.f;
            if (fl & 2) dG = 0.f;
            if (fl & 4) dB = 0.f;
            if (use_jac) {  // colour -> mean through the view direction: the forward's Jacobian times dL/dcolour
                const float m0 = jc[0] * dR + jc[3] * dG + jc[6] * dB, m1 = jc[1] * dR + jc[4] * dG + jc[7] * dB,
                            m2 = jc[2] * dR + jc[5] * dG + jc[8] * dB;
                dmean[0] += m0, dmean[1] += m1, dmean[2] += m2;
                if (POSE) camC[0] -= m0, camC[1] -= m1, camC[2] -= m2;  // dir = p - campos
            }
            if (!compact) {  // dL/dsh row = basis(dir) (x) dL/dcolour
                const float vx = mean_x - cpx, vy = mean_y - cpy, vz = mean_z - cpz;
                const float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);  // as the colour kernel normalises
                const float x = vx * inv, y = vy * inv, z = vz * inv;
#define GGRT_TERM(k, B, BX, BY, BZ)                                                               \
    {                                                                                             \
        const float b_ = (B);                                                                     \
        my[(k) * ks] = b_ * dR, my[(k) * ks + cs] = b_ * dG, my[(k) * ks + 2 * cs] = b_ * dB;     \
    }
                GGRT_SH_TERMS_0(GGRT_TERM)
                if (v.deg > 0) { GGRT_SH_TERMS_1(GGRT_TERM) }
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                if (v.deg > 1) { GGRT_SH_TERMS_2(GGRT_TERM) }
                if (v.deg > 2) { GGRT_SH_TERMS_3(GGRT_TERM) }
                if (v.deg > 3) { GGRT_SH_TERMS_4(GGRT_TERM) }
#undef GGRT_TERM
            }
        } else if (valid && !compact) {
            for (int k = 0; k < row; ++k) my[k] = 0.f;
        }
        if (push) {  // masked colour gradient (zero for culled Gaussians and clamped channels)
            if (valid) sdc[3 * threadIdx.x] = dR, sdc[3 * threadIdx.x + 1] = dG, sdc[3 * threadIdx.x + 2] = dB;
            __syncthreads();
            const int n3 = cnt * 3, n4 = n3 >> 2;
            for (int sk = 0; sk < sinks.n; ++sk) {  // every sink is a [P+1,3] buffer: local, a peer's, or multicast
                float* dstc = sinks.ptr[sk] + poff + (size_t)base * 3;  // slab offsets are multiples of 1536 B
                const bool mm = sinks.multimem != 0;
                if ((reinterpret_cast<uintptr_t>(dstc) & 15) == 0) {
                    for (int k = threadIdx.x; k < n4; k += PB_THREADS)
                        sink_store4(dstc + 4 * k, *reinterpret_cast<const float4*>(sdc + 4 * k), mm);
                    for (int k = (n4 << 2) + threadIdx.x; k < n3; k += PB_THREADS) sink_store1(dstc + k, sdc[k], mm);
                } else {
                    for (int k = threadIdx.x; k < n3; k += PB_THREADS) sink_store1(dstc + k, sdc[k], mm);
                }
            }
        }
        // write-out of the dL/dsh slab: TMA bulk store, or coalesced stores when ragged / unaligned
        if (!compact) {
            float* dst = dsh + (size_t)base * row;
            if (slab_tma) {
                fence_proxy_async();  // this thread's generic writes -> visible to the async proxy
                __syncthreads();
                if (threadIdx.x == 0) {
                    bulk_s2g(dst, smem_u32(slab), (uint32_t)nfl * 4u);
                    bulk_commit();
                }
            } else {
                __syncthreads();
                if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    float4* d4 = reinterpret_cast<float4*>(dst);
                    const float4* s4 = reinterpret_cast<const float4*>(slab);
                    const int n4 = nfl >> 2;
                    for (int k = threadIdx.x; k < n4; k += PB_THREADS) __stcs(d4 + k, s4[k]);
                    for (int k = (n4 << 2) + threadIdx.x; k < nfl; k += PB_THREADS) dst[k] = slab[k];
                } else {
                    for (int k = threadIdx.x; k < nfl; k += PB_THREADS) dst[k] = slab[k];
                }
                // (the barrier at the top of the slab loop orders these reads before the stage is written again)
            }
        } else if (push) {
            __syncthreads();  // the staged colour gradients have been read before the next slab overwrites them
        }
    } else if (valid) {
        dcolors[3 * i] = dR, dcolors[3 * i + 1] = dG, dcolors[3 * i + 2] = dB;
    }

    if (valid) {
        dmeans2D[3 * i] = g2x, dmeans2D[3 * i + 1] = g2y, dmeans2D[3 * i + 2] = 0.f;
        dopacity[i] = gop;
        if (AUX && daux) daux[i] = live ? gaux : 0.f;
        // chain rule of the scene scale: means_used = s * means, cov_used = s^2 * cov
        const float s1 = v.scale, s2 = v.scale * v.scale;
#pragma unroll
        for (int k = 0; k < 3; ++k) dmeans3D[3 * i + k] = s1 * dmean[k];
        float* dc = dcov3D + (size_t)i * v.cov_stride;
        if (v.cov_stride == 9) {  // upper triangle of the 3x3 (as the reference's triu gather does), zeros below
            dc[0] = s2 * dS[0], dc[1] = s2 * dS[1], dc[2] = s2 * dS[2];
            dc[3] = 0.f, dc[4] = s2 * dS[3], dc[5] = s2 * dS[4];
            dc[6] = 0.f, dc[7] = 0.f, dc[8] = s2 * dS[5];
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) dc[k] = s2 * dS[k];
        }
    }
    }  // slab loop
    if (POSE) {  // layout of dcam: viewmatrix [4,4] | projmatrix [4,4] | campos [3]
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            float a = camV[k], b = camM[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, d);
                b += __shfl_xor_sync(0xffffffffu, b, d);
            }
            if ((threadIdx.x & 31) == 0) {
                const int i2 = k / 3, j = k % 3;
                atomicAdd(dcam + 4 * i2 + j, a);
                atomicAdd(dcam + 16 + 4 * i2 + (j == 2 ? 3 : j), b);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = camC[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
            if ((threadIdx.x & 31) == 0) atomicAdd(dcam + 32 + k, a);
        }
    }
    if (compact && sinks.with_campos && blockIdx.x == 0 && threadIdx.x < 3) {  // row P of every sink: this view's camera centre
        for (int sk = 0; sk < sinks.n; ++sk)
            sink_store1(sinks.ptr[sk] + poff + (size_t)v.P * 3 + threadIdx.x, v.campos[threadIdx.x], sinks.multimem != 0);
    }
    if (threadIdx.x == 0) bulk_wait0();  // all bulk stores of this CTA have completed
    if (sinks.epoch != nullptr && last_cta_done(sinks.done)) {
        // every CTA's pushes and plain outputs are fenced: advance the step counter and tell the receivers
        *reinterpret_cast<volatile uint32_t*>(sinks.epoch) = *reinterpret_cast<volatile uint32_t*>(sinks.epoch) + 1u;
        for (int k = 0; k < sinks.n_arrive; ++k) signal_add(sinks.arrive[k], sinks.multimem != 0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dL/dsh writer.  dL/dsh[i] = basis(dir_i) (x) dL/dcolour_i is 88 % of the bytes the per-Gaussian backward moves
// and needs none of its arithmetic -- only the Gaussian's mean, its clamp flags and the three colour sums of the
// scratch row -- so it is a kernel of its own (40 registers, 128-thread CTAs, a store-only shared-memory ring
// drained by TMA bulk stores) that ggrt_raster_backward runs BESIDE the register-heavy, latency-bound kernel above
// on the library's side stream: together they keep HBM busy where one fused kernel could not (36.9 us fused).
// ---------------------------------------------------------------------------------------------------------------
#ifndef GGRT_SG_THREADS
#define GGRT_SG_THREADS 64
#endif
#ifndef GGRT_SG_PER_SM
#define GGRT_SG_PER_SM 4
#endif
#ifndef GGRT_SG_STAGES
#define GGRT_SG_STAGES 2
#endif
constexpr int SG_THREADS = GGRT_SG_THREADS;
constexpr int SG_STAGES = GGRT_SG_STAGES;

template <bool CMAJOR>
__global__ void __launch_bounds__(SG_THREADS)
sh_gradient_kernel(View v, const float* __restrict__ means, const int* __restrict__ radii,
                   const uint8_t* __restrict__ flags, const float* __restrict__ scratch, float* __restrict__ dsh,
                   int num_slabs) {
    extern __shared__ __align__(128) float slab_ring[];
    resolve_device_params(v);
    const int row = v.K * 3;
    const int ks = CMAJOR ? 1 : 3, cs = CMAJOR ? v.K : 1;
    const int slab_floats = SG_THREADS * row;
    const bool tma_ok = (reinterpret_cast<uintptr_t>(dsh) & 15) == 0;
    const float cpx = v.campos[0], cpy = v.campos[1], cpz = v.campos[2];

    int n_radius = 0;
    uint32_t n_flags = 0;
    float n_mean[3] = {0.f, 0.f, 0.f}, n_d[3] = {0.f, 0.f, 0.f};
    auto prefetch = [&](int sl) {
        const int i = sl * SG_THREADS + threadIdx.x;
        n_radius = 0;
        if (sl < num_slabs && i < v.P) {
            n_radius = radii[i];
            n_flags = flags[i];
            const float* gs = scratch + (size_t)i * GRAD_STRIDE;
#pragma unroll
            for (int k = 0; k < 3; ++k) n_mean[k] = means[3 * (size_t)i + k], n_d[k] = gs[G_R + k];
        }
    };
    prefetch(blockIdx.x);

    int it = 0;
    for (int sl = blockIdx.x; sl < num_slabs; sl += gridDim.x, ++it) {
        float* slab = slab_ring + (it % SG_STAGES) * slab_floats;
        const int base = sl * SG_THREADS;
        const int cnt = min(SG_THREADS, v.P - base);
        const int nfl = cnt * row;
        const bool valid = threadIdx.x < cnt, vis = valid && n_radius > 0;
        const uint32_t fl = n_flags;
        const float vx = fmul(n_mean[0], v.scale) - cpx, vy = fmul(n_mean[1], v.scale) - cpy,
                    vz = fmul(n_mean[2], v.scale) - cpz;
        const float dR = (fl & 1) ? 0.f : n_d[0], dG = (fl & 2) ? 0.f : n_d[1], dB = (fl & 4) ? 0.f : n_d[2];
        prefetch(sl + gridDim.x);
        // the bulk store issued SG_STAGES slabs ago has drained this stage
        if (threadIdx.x == 0) bulk_wait_read<SG_STAGES - 1>();
        __syncthreads();
        float* my = slab + threadIdx.x * row;
        if (vis) {
            const float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);  // as the colour kernel normalises
            const float x = vx * inv, y = vy * inv, z = vz * inv;
#define GGRT_TERM(k, B, BX, BY, BZ)                                                               \
    {                                                                                             \
        const float b_ = (B);                                                                     \
        my[(k) * ks] = b_ * dR, my[(k) * ks + cs] = b_ * dG, my[(k) * ks + 2 * cs] = b_ * dB;     \
    }
            GGRT_SH_TERMS_0(GGRT_TERM)
            if (v.deg > 0) { GGRT_SH_TERMS_1(GGRT_TERM) }
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            if (v.deg > 1) { GGRT_SH_TERMS_2(GGRT_TERM) }
            if (v.deg > 2) { GGRT_SH_TERMS_3(GGRT_TERM) }
            if (v.deg > 3) { GGRT_SH_TERMS_4(GGRT_TERM) }
#undef GGRT_TERM
        } else if (valid) {
            for (int k = 0; k < row; ++k) my[k] = 0.f;
        }
        float* dst = dsh + (size_t)base * row;
        if (tma_ok && ((nfl * 4) & 15) == 0) {
            fence_proxy_async();  // this thread's generic writes -> visible to the async proxy
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_s2g(dst, smem_u32(slab), (uint32_t)nfl * 4u);
                bulk_commit();
            }
        } else {  // ragged / unaligned slab: coalesced stores (the barrier at the top of the loop protects the stage)
            __syncthreads();
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                float4* d4 = reinterpret_cast<float4*>(dst);
                const float4* s4 = reinterpret_cast<const float4*>(slab);
                const int n4 = nfl >> 2;
                for (int k = threadIdx.x; k < n4; k += SG_THREADS) __stcs(d4 + k, s4[k]);
                for (int k = (n4 << 2) + threadIdx.x; k < nfl; k += SG_THREADS) dst[k] = slab[k];
            } else {
                for (int k = threadIdx.x; k < nfl; k += SG_THREADS) dst[k] = slab[k];
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait0();  // all bulk stores of this CTA have completed
}

void launch_preprocess_backward(const View& v, const float* means, const float* cov3d, const float* shs,
                                const int* radii, GeomPtrs g, const float* scratch, float* dmeans2D, float* dopacity,
                                float* dmeans3D, float* dcov3D, float* dsh, float* dcolors, float* daux, float* dcam,
                                const ColorSinks& sinks, cudaStream_t s, cudaStream_t sh_stream) {
    if (v.P == 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#ifndef GGRT_PB_FUSED_DSH
    if (shs != nullptr && dsh != nullptr) {
        // dL/dsh goes through its own streaming kernel (on `sh_stream`, which the caller has ordered after the render
        // backward and joins afterwards -- or on `s` itself); the kernel below then runs without any SH work
        const size_t smem = (size_t)SG_STAGES * SG_THREADS * v.K * 3 * sizeof(float);
        const int num_slabs = (v.P + SG_THREADS - 1) / SG_THREADS;
        const int per_sm = max(1, min(GGRT_SG_PER_SM, (int)((200 * 1024) / (smem + 1024))));
        const int grid = min(num_slabs, per_sm * sms);
        cudaStream_t st = sh_stream ? sh_stream : s;
        if (v.sh_ks == 1 && v.K > 1) {
            if (smem > 32 * 1024)
                cudaFuncSetAttribute(sh_gradient_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            sh_gradient_kernel<true><<<grid, SG_THREADS, smem, st>>>(v, means, radii, g.flags, scratch, dsh, num_slabs);
        } else {
            if (smem > 32 * 1024)
                cudaFuncSetAttribute(sh_gradient_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            sh_gradient_kernel<false><<<grid, SG_THREADS, smem, st>>>(v, means, radii, g.flags, scratch, dsh, num_slabs);
        }
        dsh = nullptr;  // (shs != NULL, no dsh, no sinks: the "compact" code path with nothing to push)
    }
#endif
    const size_t smem = (shs && dsh) ? (size_t)PB_STAGES * PB_THREADS * v.K * 3 * sizeof(float) : 0;  // dL/dsh ring
    const int num_slabs = (v.P + PB_THREADS - 1) / PB_THREADS;
    // persistent CTAs: as many per SM as shared memory AND the register budget of __launch_bounds__ allow
    // (without a ring: GGRT_PB_SPLIT_BLOCKS one-warp CTAs per SM.  Measured at C2 with the dL/dsh writer beside it, whole
    // step: 9 CTAs x 166 registers 0.2923 ms, 13 x 128 0.2921, 16 x 128 0.2911, 17 x 96 0.2948)
    const int per_sm = dcam ? 10 : smem ? max(1, min(GGRT_PB_MINBLOCKS, (int)((220 * 1024) / (smem + 1024)))) : GGRT_PB_SPLIT_BLOCKS;
    const int grid = min(num_slabs, per_sm * sms);
#define GGRT_LAUNCH_PB(AX, CM, PO)                                                                                      \
    {                                                                                                                   \
        if (smem > 32 * 1024)                                                                                           \
            cudaFuncSetAttribute(preprocess_backward_kernel<AX, CM, PO>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem);                                                                            \
        launch_chain(preprocess_backward_kernel<AX, CM, PO>, dim3(grid), dim3(PB_THREADS), smem, s, v, means, cov3d, shs,  \
                     radii, (const uint8_t*)g.flags, (const float*)g.jac, g.jac_plane, scratch, dmeans2D, dopacity,     \
                     dmeans3D, dcov3D, dsh, dcolors, daux, dcam, sinks, num_slabs);                                     \
    }
    const bool cm = v.sh_ks == 1 && v.K > 1;
    if (dcam) {  // camera gradients are rare: one instantiation per SH layout, aux always compiled in
        if (cm) GGRT_LAUNCH_PB(true, true, true) else GGRT_LAUNCH_PB(true, false, true)
    } else if (daux || v.aux_mode) {
        if (cm) GGRT_LAUNCH_PB(true, true, false) else GGRT_LAUNCH_PB(true, false, false)
    } else {
        if (cm) GGRT_LAUNCH_PB(false, true, false) else GGRT_LAUNCH_PB(false, false, false)
    }
#undef GGRT_LAUNCH_PB
}

}  // namespace ggrt
