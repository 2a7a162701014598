// Forward per-Gaussian kernels: geometry (A.1 cull / project / cov2D / radius / tile
// rect + per-tile pair counting), tile scan, SH colour evaluation, mark_visible.
// Replaces upstream preprocessCUDA + the scan of tiles_touched.
#include "common.cuh"
#include "tma.cuh"

namespace ggrt {

#ifndef GGRT_GEO_THREADS
#define GGRT_GEO_THREADS 256
#endif
constexpr int GEO_THREADS = GGRT_GEO_THREADS;
// x / 16 as x * (1/16): TILE is a power of two, so both are the correctly rounded value of the same real number --
// bit-identical to the oracle's division for every float (incl. subnormal results), at 1 instead of ~11 instructions
#ifndef GGRT_GEO_DIV_BY_MUL
#define GGRT_GEO_DIV_BY_MUL 0
#endif
static_assert((TILE & (TILE - 1)) == 0, "the tile edge must be a power of two");
__device__ __forceinline__ float div_tile(float x) {
#if GGRT_GEO_DIV_BY_MUL
    return fmul(x, 1.0f / (float)TILE);
#else
    return fdiv(x, (float)TILE);
#endif
}
#ifndef GGRT_GEO_MINBLOCKS
#define GGRT_GEO_MINBLOCKS 4
#endif
#ifndef GGRT_COLOR_THREADS
#define GGRT_COLOR_THREADS 128
#endif
#ifndef GGRT_COLOR_STAGES
#define GGRT_COLOR_STAGES 2
#endif
constexpr int COLOR_THREADS = GGRT_COLOR_THREADS;

// ---------------------------------------------------------------------------------------
// geometry: one thread per Gaussian.  Reads 40 B, writes 2 records + rect + radius.
// Pair counting goes straight into the per-tile counters (no per-Gaussian prefix sum:
// tile segments are allocated per tile, see binning.cu).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEO_THREADS, GGRT_GEO_MINBLOCKS)
geometry_kernel(View v, const float* __restrict__ means, const float* __restrict__ cov3d,
                const float* __restrict__ opac, int* __restrict__ radii, GeomPtrs g, uint32_t* __restrict__ counts) {
    __shared__ float sV[16], sM[16];
    if (threadIdx.x < 16) sV[threadIdx.x] = v.view[threadIdx.x];
    else if (threadIdx.x < 32) sM[threadIdx.x - 16] = v.proj[threadIdx.x - 16];
    resolve_device_params(v);
    __syncthreads();
    const int i = blockIdx.x * GEO_THREADS + threadIdx.x;
    if (i >= v.P) return;

    const float px = fmul(means[3 * i], v.scale), py = fmul(means[3 * i + 1], v.scale), pz = fmul(means[3 * i + 2], v.scale);
    float cv[6];
    load_cov6(v, cov3d, i, cv);
    const float o = opac[i];

    int radius = 0;
    uint32_t tiles = 0;
    ushort4 rect = make_ushort4(0, 0, 0, 0);
    float4 r0 = make_float4(0.f, 0.f, -1.0f, 0.f);
    float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 rk = make_uint4(0u, 0u, 0u, 0u);

    Geo q;
    if (geometry(v, sV, sM, px, py, pz, cv, q)) {
        const float dinv = fdiv(1.0f, q.det);
        const float cA = fmul(q.c, dinv), cB = fmul(-q.b, dinv), cC = fmul(q.a, dinv);
        const float mid = fmul(0.5f, fadd(q.a, q.c));
        const float sq = fsqrt(fmaxf(0.1f, fsub(fmul(mid, mid), q.det)));
        const float l1 = fadd(mid, sq), l2 = fsub(mid, sq);
        const float radf = fminf(ceilf(fmul(3.0f, fsqrt(fmaxf(l1, l2)))), RADIUS_CAP);
        const float ndcx = fmul(q.hx, q.pw), ndcy = fmul(q.hy, q.pw);
        const float pxx = fmul(fsub(fmul(fadd(ndcx, 1.0f), (float)v.W), 1.0f), 0.5f);
        const float pxy = fmul(fsub(fmul(fadd(ndcy, 1.0f), (float)v.H), 1.0f), 0.5f);
        const int x0 = min(v.gx, max(0, f2i_sat(div_tile(fsub(pxx, radf)))));
        const int y0 = min(v.gy, max(0, f2i_sat(div_tile(fsub(pxy, radf)))));
        const int x1 = min(v.gx, max(0, f2i_sat(div_tile(fadd(fadd(pxx, radf), (float)(TILE - 1))))));
        const int y1 = min(v.gy, max(0, f2i_sat(div_tile(fadd(fadd(pxy, radf), (float)(TILE - 1))))));
        const int area = (x1 - x0) * (y1 - y0);
        if (area > 0) {
            // (the counting atomics go first: their round trips overlap the conic / threshold arithmetic below)
            if (area <= RANKED_TILES) {
                // small splat (the common case: 1, 2 or 2x2 tiles): count with RETURNING atomics, all of them in flight
                // together, and keep the ranks -- emit then needs no atomic for these pairs
                const int sub = i & (SUB_LANES - 1), w = x1 - x0;
                uint32_t r[RANKED_TILES];
#pragma unroll
                for (int u = 0; u < RANKED_TILES; ++u) {
                    r[u] = 0u;
                    if (u < area) r[u] = atomicAdd(&counts[(((y0 + u / w) * v.gx + x0 + u % w) << SUBS_LOG2) + sub], 1u);
                }
                rk = make_uint4(r[0], r[1], r[2], r[3]);
            } else {
                const int sub = SUB_LANES + (i & (SUB_LANES - 1));  // the bank emit allocates from with atomics
                for (int y = y0; y < y1; ++y)
                    for (int x = x0; x < x1; ++x) atomicAdd(&counts[((y * v.gx + x) << SUBS_LOG2) + sub], 1u);
            }
            radius = (int)radf;
            tiles = (uint32_t)area;
            rect = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
            // {alpha >= 1/255} <=> q(d) = A dx^2 + 2B dx dy + C dy^2 <= tau = 2 ln(255 o).  The render kernels cull
            // (warp pixel block, Gaussian) pairs with an exact ellipse-vs-rectangle test against tau, slightly
            // inflated so that float rounding of the per-pixel power can never contradict the cull.
            float tau_c = -1.0f;  // opacity below 1/255: never visible
            const float tau = 2.0f * logf(255.0f * o);
            if (tau > 0.0f) tau_c = tau * 1.001f + 0.02f;
            r0 = make_float4(pxx, pxy, tau_c, q.tz);
            r1 = make_float4(cA, cB, cC, o);
        }
    }
    radii[i] = radius;
    g.tiles[i] = tiles;
    g.rect[i] = rect;
    g.rec0[i] = r0;
    g.rec1[i] = r1;
    g.ranks[i] = rk;  // last: the only store that has to wait for the atomics' results (issue is in order)
}

// ---------------------------------------------------------------------------------------
// scan_tiles: exclusive scan of the T*SUBS pair counters (L2 resident).  Writes the
// per-(tile, sub-counter) segment starts that `emit` consumes as allocation cursors, the
// per-tile starts for sort / render, and reports {N, max pairs in a tile}.
// One tile per thread, SCAN_BLOCK tiles per CTA; CTAs chain through a single-pass look-back:
// each publishes {flag, max, total} and sums the totals of the CTAs before it (they are
// scheduled in index order and never wait on later ones, so the spin cannot deadlock).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tiles_kernel(int T, const uint32_t* __restrict__ counts,
                                                                uint32_t* __restrict__ starts,
                                                                uint32_t* __restrict__ sub_starts,
                                                                unsigned long long* partials,
                                                                uint32_t* __restrict__ header,
                                                                volatile uint32_t* counts_host) {
    // Running sums are 64-bit: the pair lists are indexed with 32 bits everywhere downstream, so a total of 2^32 or
    // more cannot be rendered -- but it must be DETECTED (header[2] / counts_host[2] carry the high word, the host
    // raises GGRT_ERR_UNSUPPORTED) instead of wrapping silently.
    __shared__ unsigned long long warp_sums[SCAN_BLOCK / 32];
    __shared__ uint32_t warp_max[SCAN_BLOCK / 32];
    __shared__ unsigned long long prefix_s;
    __shared__ uint32_t gmax_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x * SCAN_BLOCK + tid;
    pdl_enter();
    const uint4* c4 = reinterpret_cast<const uint4*>(counts) + (size_t)t * (SUBS / 4);
    uint4 c[SUBS / 4];
    unsigned long long ts = 0;
#pragma unroll
    for (int q = 0; q < SUBS / 4; ++q) {
        c[q] = (t < T) ? c4[q] : make_uint4(0u, 0u, 0u, 0u);
        ts += (unsigned long long)c[q].x + c[q].y + c[q].z + c[q].w;
    }
    unsigned long long x = ts;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, (uint32_t)min(ts, 0x7fffffffull));
    if (lane == 31) warp_sums[warp] = x, warp_max[warp] = wmax;
    if (tid == 0) prefix_s = 0, gmax_s = 0;
    __syncthreads();
    unsigned long long block_total = 0, warp_off = 0;
    uint32_t block_max = 0;
#pragma unroll
    for (int w = 0; w < SCAN_BLOCK / 32; ++w) {
        if (w < warp) warp_off += warp_sums[w];
        block_total += warp_sums[w];
        block_max = max(block_max, warp_max[w]);
    }
    // publish this CTA's aggregate: {bit 63 flag | 63 bits total} in partials[b], the maximum in partials[nb + b]
    const int nb = gridDim.x;
    if (tid == 0) {
        *reinterpret_cast<volatile unsigned long long*>(&partials[nb + blockIdx.x]) = block_max;
        __threadfence();
        atomicExch(&partials[blockIdx.x], (1ull << 63) | (block_total & 0x7fffffffffffffffull));
    }
    // look back: sum the totals (and fold the maxima) of all earlier CTAs
    unsigned long long psum = 0;
    uint32_t pmax = 0;
    for (int bk = tid; bk < (int)blockIdx.x; bk += SCAN_BLOCK) {
        unsigned long long v;
        do {
            v = *reinterpret_cast<volatile unsigned long long*>(&partials[bk]);
        } while ((v >> 63) == 0);
        __threadfence();
        psum += v & 0x7fffffffffffffffull;
        pmax = max(pmax, (uint32_t) * reinterpret_cast<volatile unsigned long long*>(&partials[nb + bk]));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, d);
    pmax = __reduce_max_sync(0xffffffffu, pmax);
    if (lane == 0 && (psum | pmax)) atomicAdd(&prefix_s, psum), atomicMax(&gmax_s, pmax);
    __syncthreads();
    unsigned long long run64 = prefix_s + warp_off + x - ts;
    if (t < T) {
        uint32_t run = (uint32_t)run64;
        starts[t] = run;
        uint4* s4 = reinterpret_cast<uint4*>(sub_starts) + (size_t)t * (SUBS / 4);
#pragma unroll
        for (int q = 0; q < SUBS / 4; ++q) {
            uint4 o;
            o.x = run; run += c[q].x;
            o.y = run; run += c[q].y;
            o.z = run; run += c[q].z;
            o.w = run; run += c[q].w;
            s4[q] = o;
        }
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        const unsigned long long total = prefix_s + block_total;
        const uint32_t lo = (uint32_t)total, hi = (uint32_t)(total >> 32), mx = max(gmax_s, block_max);
        starts[T] = lo;
        header[0] = lo;
        header[1] = mx;
        header[2] = hi;  // non-zero: more than 2^32 - 1 pairs (unsupported)
        header[3] = 0u;
        if (counts_host != nullptr) {  // mapped pinned host memory: the host polls / waits on the following event
            counts_host[0] = lo;
            counts_host[1] = mx;
            counts_host[2] = hi;
            counts_host[3] = 0u;
#ifndef GGRT_SCAN_NO_SYSFENCE
            __threadfence_system();
#endif
        }
    }
}

// ---------------------------------------------------------------------------------------
// colour: SH degree 0..4 -> RGB (+0.5, clamp at 0, remember clamp bits).  HBM-bound: 12 K
// bytes per Gaussian (300 B at degree 4) are read exactly once -- in the whole step: while the row
// is on chip the kernel also contracts it with the basis derivatives into d(colour)/d(mean)
// (GeomPtrs::jac, 3x3 floats), which is all the backward needs the coefficients for.
//
// Persistent CTAs (2 per SM); each walks slabs of COLOR_THREADS Gaussians.  A slab's SH block
// is contiguous in HBM (COLOR_THREADS x K x 3 floats) and is pulled into a 2-stage shared
// memory ring by the TMA engine (cp.async.bulk + mbarrier): one elected thread issues the
// copy of slab i+2 while all threads evaluate slab i.  Each thread then walks its own row
// (odd stride: no bank conflicts for K = 1, 9, 25).  Ragged / unaligned slabs fall back to
// cooperative 16-byte (or scalar) loads into the same buffer.
// ---------------------------------------------------------------------------------------
constexpr int COLOR_STAGES = GGRT_COLOR_STAGES;

__device__ __forceinline__ void fill_slab_generic(float* slab, const float* src, int nfl, int nthreads) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(slab);
        const int n4 = nfl >> 2;
        for (int k = threadIdx.x; k < n4; k += nthreads) d4[k] = __ldcs(s4 + k);
        for (int k = (n4 << 2) + threadIdx.x; k < nfl; k += nthreads) slab[k] = src[k];
    } else {
        for (int k = threadIdx.x; k < nfl; k += nthreads) slab[k] = src[k];
    }
}

template <int DEG, bool CMAJOR, bool JAC>
__global__ void __launch_bounds__(COLOR_THREADS, 3)
color_kernel(View v, const float* __restrict__ means, const float* __restrict__ shs, const float* __restrict__ colors,
             const float* __restrict__ aux, const int* __restrict__ radii, GeomPtrs g, int num_slabs) {
    extern __shared__ __align__(128) float slab_ring[];
    __shared__ __align__(8) unsigned long long full_bar[COLOR_STAGES];
    constexpr int KK = (DEG + 1) * (DEG + 1);
    constexpr int row = KK * 3;
    constexpr int ks = CMAJOR ? 1 : 3, cs = CMAJOR ? KK : 1;  // SH element (k, c) at k*ks + c*cs of the row
    resolve_device_params(v);
    const int slab_floats = COLOR_THREADS * row;
    const bool tma_ok = shs != nullptr && (reinterpret_cast<uintptr_t>(shs) & 15) == 0;  // slab bases are 16 B multiples

    if (shs != nullptr) {
        if (threadIdx.x == 0) {
            for (int st = 0; st < COLOR_STAGES; ++st) mbar_init(smem_u32(&full_bar[st]), 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0 && tma_ok) {  // prologue: fill the ring
            for (int st = 0; st < COLOR_STAGES; ++st) {
                const int sl = blockIdx.x + st * gridDim.x;
                if (sl >= num_slabs) break;
                const int cnt = min(COLOR_THREADS, v.P - sl * COLOR_THREADS);
                const uint32_t bytes = (uint32_t)cnt * row * 4u;
                if ((bytes & 15u) == 0) {
                    mbar_expect_tx(smem_u32(&full_bar[st]), bytes);
                    bulk_g2s(smem_u32(slab_ring + st * slab_floats), shs + (size_t)sl * slab_floats, bytes,
                             smem_u32(&full_bar[st]));
                }
            }
        }
    }

    // per-Gaussian scalars of the NEXT slab are prefetched into registers while the current one is evaluated,
    // so their DRAM latency is not serialised with the slab wait
    int n_radius = 0;
    float n_mx = 0.f, n_my = 0.f, n_mz = 0.f, n_depth = 0.f;
    auto prefetch = [&](int sl) {
        const int i = sl * COLOR_THREADS + threadIdx.x;
        n_radius = 0;
        if (sl < num_slabs && i < v.P) {
            n_radius = radii[i];
            n_mx = means[3 * i], n_my = means[3 * i + 1], n_mz = means[3 * i + 2];  // raw: scaled when consumed
            n_depth = aux ? aux[i] : g.rec0[i].w;  // 4th blended channel: caller's aux or the view depth ...
            if (v.aux_mode == 1) n_depth = ggrt_depth_channel(n_depth, v.scale);  // ... or GGRt's depth channel of it
        }
    };
    prefetch(blockIdx.x);
    const float cpx = v.campos[0], cpy = v.campos[1], cpz = v.campos[2];

    int it = 0;
    for (int sl = blockIdx.x; sl < num_slabs; sl += gridDim.x, ++it) {
        const int st = it % COLOR_STAGES;
        const uint32_t parity = (uint32_t)(it / COLOR_STAGES) & 1u;
        const int base = sl * COLOR_THREADS;
        const int cnt = min(COLOR_THREADS, v.P - base);
        const int i = base + threadIdx.x;
        const bool valid = threadIdx.x < cnt;
        const bool vis = valid && n_radius > 0;
        const float mx = fmul(n_mx, v.scale), my_ = fmul(n_my, v.scale), mz = fmul(n_mz, v.scale), depth = n_depth;
        prefetch(sl + gridDim.x);
        float* slab = slab_ring + st * slab_floats;
        if (shs != nullptr) {
            const uint32_t bytes = (uint32_t)cnt * row * 4u;
            if (tma_ok && (bytes & 15u) == 0) {
                mbar_wait(smem_u32(&full_bar[st]), parity);
            } else {
                fill_slab_generic(slab, shs + (size_t)base * row, cnt * row, COLOR_THREADS);
                __syncthreads();
            }
        }
        if (valid) {
            float rgb[3] = {0.f, 0.f, 0.f};
            uint8_t flags = 0;
            if (vis) {
                if (shs == nullptr) {
                    rgb[0] = colors[3 * i], rgb[1] = colors[3 * i + 1], rgb[2] = colors[3 * i + 2];
                } else {
                    const float vx = mx - cpx, vy = my_ - cpy, vz = mz - cpz;
                    const float len2 = vx * vx + vy * vy + vz * vz;
                    const float inv = 1.0f / sqrtf(len2);
                    const float x = vx * inv, y = vy * inv, z = vz * inv;
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    const float* my = slab + threadIdx.x * row;
                    rgb[0] = rgb[1] = rgb[2] = 0.5f;
                    // d(colour c)/d(direction) accumulates beside the colour: the SH row is in shared memory here, and
                    // with this 3x3 Jacobian stored the backward pass never reads the 12K-byte row again
                    float jx[3] = {0.f, 0.f, 0.f}, jy[3] = {0.f, 0.f, 0.f}, jz[3] = {0.f, 0.f, 0.f};
#define GGRT_COLOR_TERM(k, B, BX, BY, BZ)                                                       \
    {                                                                                           \
        const float b_ = (B);                                                                   \
        _Pragma("unroll") for (int c = 0; c < 3; ++c) {                                         \
            const float s_ = my[(k) * ks + c * cs];                                             \
            rgb[c] = fmaf(b_, s_, rgb[c]);                                                      \
            if (JAC) {                                                                          \
                sh_fma(BX, s_, jx[c]);                                                            \
                sh_fma(BY, s_, jy[c]);                                                            \
                sh_fma(BZ, s_, jz[c]);                                                            \
            }                                                                                   \
        }                                                                                       \
    }
                    GGRT_SH_TERMS_0(GGRT_COLOR_TERM)
                    if (DEG > 0) { GGRT_SH_TERMS_1(GGRT_COLOR_TERM) }
                    if (DEG > 1) { GGRT_SH_TERMS_2(GGRT_COLOR_TERM) }
                    if (DEG > 2) { GGRT_SH_TERMS_3(GGRT_COLOR_TERM) }
                    if (DEG > 3) { GGRT_SH_TERMS_4(GGRT_COLOR_TERM) }
#undef GGRT_COLOR_TERM
                    if (JAC) {  // through normalize(): (I |v|^2 - v v^T) / |v|^3, one row of the Jacobian per channel
                        const float inv3 = inv * inv * inv;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float dot = vx * jx[c] + vy * jy[c] + vz * jz[c];
                            g.jac[(size_t)(3 * c + 0) * g.jac_plane + i] = (len2 * jx[c] - vx * dot) * inv3;
                            g.jac[(size_t)(3 * c + 1) * g.jac_plane + i] = (len2 * jy[c] - vy * dot) * inv3;
                            g.jac[(size_t)(3 * c + 2) * g.jac_plane + i] = (len2 * jz[c] - vz * dot) * inv3;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (rgb[c] < 0.0f) flags |= (uint8_t)(1u << c);
                        rgb[c] = fmaxf(rgb[c], 0.0f);
                    }
                }
                g.rec2[i] = make_float4(rgb[0], rgb[1], rgb[2], depth);
            }
            g.flags[i] = flags;
        }
        if (shs != nullptr) {
            __syncthreads();  // every thread is done reading this stage
            const int nsl = sl + COLOR_STAGES * gridDim.x;
            if (threadIdx.x == 0 && tma_ok && nsl < num_slabs) {
                const int ncnt = min(COLOR_THREADS, v.P - nsl * COLOR_THREADS);
                const uint32_t nbytes = (uint32_t)ncnt * row * 4u;
                if ((nbytes & 15u) == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(smem_u32(&full_bar[st]), nbytes);
                    bulk_g2s(smem_u32(slab), shs + (size_t)nsl * slab_floats, nbytes, smem_u32(&full_bar[st]));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// camera_setup: one thread per view, float32 arithmetic in the order of the reference glue where that matters for
// the result (projection entries, tan(fov/2)); the rigid inverse is evaluated in double.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void inv3x3(const float* K, double* o) {
    const double a = K[0], b = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C, r = 1.0 / det;
    o[0] = A * r, o[1] = -(b * i - c * h) * r, o[2] = (b * f - c * e) * r;
    o[3] = B * r, o[4] = (a * i - c * g) * r, o[5] = -(a * f - c * d) * r;
    o[6] = C * r, o[7] = -(a * h - b * g) * r, o[8] = (a * e - b * d) * r;
}

__global__ void camera_setup_kernel(int n, const float* __restrict__ extr, const float* __restrict__ intr,
                                    const float* __restrict__ near, const float* __restrict__ far, int scale_invariant,
                                    float* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const float* E = extr + 16 * v;
    const float* K = intr + 9 * v;
    float* o = out + (size_t)GGRT_CAMERA_FLOATS * v;
    const float scale = scale_invariant ? fdiv(1.0f, near[v]) : 1.0f;
    const float nr = fmul(near[v], scale), fr = fmul(far[v], scale);
    // camera-to-world with the translation rescaled (cuda_splatting.py:68-69); its inverse, transposed = viewmatrix
    double M[16];
    for (int k = 0; k < 16; ++k) M[k] = E[k];
    for (int r = 0; r < 3; ++r) M[4 * r + 3] = (double)fmul(E[4 * r + 3], scale);
    // general 4x4 inverse by cofactors (the reference calls .inverse(); poses are rigid but nothing here assumes it)
    double inv[16];
    {
        const double* m = M;
        inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
        inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
        inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
        inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
        inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
        inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
        inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
        inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
        inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
        inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
        inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
        inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
        inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
        inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
        inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
        inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
        const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12], r = 1.0 / det;
        for (int k = 0; k < 16; ++k) inv[k] *= r;
    }
    float V[16];  // viewmatrix = inverse(extrinsics)^T, float32 as the reference holds it
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) V[4 * r + c] = (float)inv[4 * c + r];
    // projection (get_projection_matrix, cuda_splatting.py:18-46) from the intrinsics of VIEW 0 (reference quirk)
    const float* K0 = intr;
    float Pm[16];
    for (int k = 0; k < 16; ++k) Pm[k] = 0.0f;
    Pm[0] = fmul(fmul(2.0f, nr), K0[0]);
    Pm[5] = fmul(fmul(2.0f, nr), K0[4]);
    Pm[2] = fsub(fmul(2.0f, K0[2]), 1.0f);
    Pm[6] = fsub(fmul(2.0f, K0[5]), 1.0f);
    Pm[14] = 1.0f;
    Pm[10] = fdiv(fr, fsub(fr, nr));
    Pm[11] = fdiv(-fmul(fr, nr), fsub(fr, nr));
    // full = view @ projection^T (row-vector convention), float32 accumulation in index order as torch's matmul of
    // a [4,4] pair does for these sizes (differences are at the last ulp and inside the parity tolerance)
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) acc = fmaf(V[4 * r + k], Pm[4 * c + k], acc);
            o[16 + 4 * r + c] = acc;
        }
    for (int k = 0; k < 16; ++k) o[k] = V[k];
    for (int r = 0; r < 3; ++r) o[32 + r] = fmul(E[4 * r + 3], scale);  // campos = rescaled camera-to-world translation
    // get_fov (ggrt/geometry/projection.py:233-247) of THIS view: angle between the un-projected mid-points of
    // opposite image edges
    double Ki[9];
    inv3x3(K, Ki);
    const double px[4][3] = {{0, 0.5, 1}, {1, 0.5, 1}, {0.5, 0, 1}, {0.5, 1, 1}};
    double ray[4][3];
    for (int p = 0; p < 4; ++p) {
        double nrm = 0;
        for (int r = 0; r < 3; ++r) {
            ray[p][r] = Ki[3 * r] * px[p][0] + Ki[3 * r + 1] * px[p][1] + Ki[3 * r + 2] * px[p][2];
            nrm += ray[p][r] * ray[p][r];
        }
        nrm = sqrt(nrm);
        for (int r = 0; r < 3; ++r) ray[p][r] /= nrm;
    }
    const double cx_ = ray[0][0] * ray[1][0] + ray[0][1] * ray[1][1] + ray[0][2] * ray[1][2];
    const double cy_ = ray[2][0] * ray[3][0] + ray[2][1] * ray[3][1] + ray[2][2] * ray[3][2];
    const float fovx = (float)acos(fmin(1.0, fmax(-1.0, cx_))), fovy = (float)acos(fmin(1.0, fmax(-1.0, cy_)));
    o[35] = tanf(fmul(0.5f, fovx));
    o[36] = tanf(fmul(0.5f, fovy));
    o[37] = scale;
    o[38] = nr, o[39] = fr;
    for (int k = 40; k < GGRT_CAMERA_FLOATS; ++k) o[k] = 0.0f;
}

void launch_camera_setup(int n, const float* extr, const float* intr, const float* near, const float* far,
                         int scale_invariant, float* out, cudaStream_t s) {
    if (n <= 0) return;
    camera_setup_kernel<<<(n + 31) / 32, 32, 0, s>>>(n, extr, intr, near, far, scale_invariant, out);
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means, const float* __restrict__ V,
                                    uint8_t* __restrict__ present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float tz = fadd(dot3(V[2], means[3 * i], V[6], means[3 * i + 1], V[10], means[3 * i + 2]), V[14]);
    present[i] = (uint8_t)(tz > NEAR_CULL);
}

void launch_geometry(const View& v, const float* means, const float* cov3d, const float* opac, int* radii, GeomPtrs g,
                     ImagePtrs im, cudaStream_t s) {
    if (v.P == 0) return;
    geometry_kernel<<<(v.P + GEO_THREADS - 1) / GEO_THREADS, GEO_THREADS, 0, s>>>(v, means, cov3d, opac, radii, g,
                                                                                   im.counts);
}

void launch_scan_tiles(const View& v, ImagePtrs im, uint32_t* counts_host, cudaStream_t s) {
    const int T = v.gx * v.gy;
    launch_chain(scan_tiles_kernel, dim3((T + SCAN_BLOCK - 1) / SCAN_BLOCK), dim3(SCAN_BLOCK), 0, s, T, im.counts, im.starts,
                 im.cursor, im.partials, im.header, counts_host);
}

void launch_color(const View& v, const float* means, const float* shs, const float* colors, const float* aux,
                  const int* radii, GeomPtrs g, cudaStream_t s) {
    if (v.P == 0) return;
    const size_t smem = shs ? (size_t)COLOR_STAGES * COLOR_THREADS * v.K * 3 * sizeof(float) : 0;
    const int num_slabs = (v.P + COLOR_THREADS - 1) / COLOR_THREADS;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = smem ? max(1, min(8, (int)((220 * 1024) / (smem + 1024)))) : 8;
    const int grid = min(num_slabs, per_sm * sms);  // persistent CTAs, each streams slabs through its ring
    // (the direction Jacobian for the backward is stored whenever there is an SH table of degree > 0 to differentiate)
#define GGRT_LAUNCH_COLOR2(D, CM)                                                                                   \
    {                                                                                                               \
        constexpr bool JC = (D) > 0;                                                                                \
        if (smem > 32 * 1024)                                                                                       \
            cudaFuncSetAttribute(color_kernel<D, CM, JC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (shs != nullptr)                                                                                         \
            color_kernel<D, CM, JC><<<grid, COLOR_THREADS, smem, s>>>(v, means, shs, colors, aux, radii, g, num_slabs); \
        else                                                                                                        \
            color_kernel<0, false, false><<<grid, COLOR_THREADS, smem, s>>>(v, means, shs, colors, aux, radii, g, num_slabs); \
    }
#define GGRT_LAUNCH_COLOR(D)                                                                                        \
    case D:                                                                                                         \
        if (v.sh_ks == 1 && D > 0) GGRT_LAUNCH_COLOR2(D, true) else GGRT_LAUNCH_COLOR2(D, false)                     \
        break;
    switch (v.deg) {
        GGRT_LAUNCH_COLOR(0)
        GGRT_LAUNCH_COLOR(1)
        GGRT_LAUNCH_COLOR(2)
        GGRT_LAUNCH_COLOR(3)
        GGRT_LAUNCH_COLOR(4)
    }
#undef GGRT_LAUNCH_COLOR
#undef GGRT_LAUNCH_COLOR2
}

void launch_mark_visible(int P, const float* means, const float* view, uint8_t* present, cudaStream_t s) {
    if (P == 0) return;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means, view, present);
}

}  // namespace ggrt
