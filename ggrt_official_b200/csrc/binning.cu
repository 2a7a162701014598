// Tile binning: per-tile segment allocation, key emission and an in-tile sort.
// Replaces upstream duplicateWithKeys + 64-bit DeviceRadixSort + identifyTileRanges
// (SURVEY.md 8a rows a8-a10).  The resulting order inside each tile is the reference
// order: ascending (view depth bits, Gaussian index); tile segments are laid out in
// ascending tile id, so the concatenation equals the reference's globally sorted list.
//
// Instead of sorting N (u64,u32) pairs globally (5-6 radix passes over 24 B x N), each
// tile owns the contiguous segment [starts[t], starts[t+1]) and sorts its few hundred
// (depth_bits << 32 | idx) keys in shared memory with one CTA.  Keys are unique, so the
// result is deterministic although emission order is not.
#include "common.cuh"

namespace ggrt {

#ifndef GGRT_SORT_Q
#define GGRT_SORT_Q 1
#endif
constexpr int EMIT_THREADS = 256;
constexpr int SORT_THREADS = 256;

__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(View v, GeomPtrs g, uint32_t* __restrict__ cursor, unsigned long long* __restrict__ keys,
            uint32_t capacity) {
    pdl_enter();
    const int i = blockIdx.x * EMIT_THREADS + threadIdx.x;
    if (i >= v.P) return;
    // all three per-Gaussian loads are issued together (a culled Gaussian has an empty rect: no separate look at
    // tiles_touched, which would put one more dependent round trip in front of everything else)
    const ushort4 r = g.rect[i];
    const uint4 rk = g.ranks[i];
    const unsigned long long key = ((unsigned long long)__float_as_uint(g.rec0[i].w) << 32) | (uint32_t)i;
    const int w = r.z - r.x, n = w * (r.w - r.y);
    if (n == 0) return;
    if (n <= RANKED_TILES) {
        // small splat: the geometry kernel's counting atomics already returned this pair's rank inside its
        // (tile, sub-counter) segment -- slot = segment start + rank, no atomic, nothing to wait for but two loads
        const int sub = i & (SUB_LANES - 1);
        const uint32_t rank[RANKED_TILES] = {rk.x, rk.y, rk.z, rk.w};
        uint32_t slot[RANKED_TILES];
#pragma unroll
        for (int u = 0; u < RANKED_TILES; ++u) {
            slot[u] = 0xffffffffu;
            if (u < n) slot[u] = cursor[(((r.y + u / w) * v.gx + r.x + u % w) << SUBS_LOG2) + sub] + rank[u];
        }
#pragma unroll
        for (int u = 0; u < RANKED_TILES; ++u)
            if (slot[u] < capacity) keys[slot[u]] = key;  // a too-small (speculative) buffer is detected and redone by the host
        return;
    }
    // Larger splats: the (tile, sub-counter) segment start of the second counter bank doubles as its allocation cursor,
    // one returning atomic per pair.  The claims are issued four at a time and only then the dependent stores: four round
    // trips in flight per thread instead of one.
    const int sub = SUB_LANES + (i & (SUB_LANES - 1));
    for (int t0 = 0; t0 < n; t0 += 4) {
        uint32_t slot[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u;
            slot[u] = 0xffffffffu;
            if (t < n) {
                const int y = r.y + t / w, x = r.x + t % w;
                slot[u] = atomicAdd(&cursor[((y * v.gx + x) << SUBS_LOG2) + sub], 1u);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (slot[u] < capacity) keys[slot[u]] = key;  // a too-small (speculative) buffer is detected and redone by the host
    }
}

__device__ __forceinline__ void cmpxchg(unsigned long long* a, uint32_t i, uint32_t l) {
    const unsigned long long x = a[i], y = a[l];
    if (x > y) {
        a[i] = y;
        a[l] = x;
    }
}

// Ascending-only bitonic network ("flip" formulation): every compare-exchange puts the
// minimum at the lower index, so indices >= n behave as +inf padding and are skipped.
__device__ __forceinline__ void bitonic_sort(unsigned long long* a, uint32_t n, uint32_t nthreads = SORT_THREADS) {
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    const uint32_t half = n2 >> 1;
    for (uint32_t lk = 1; (1u << lk) <= n2; ++lk) {  // k = 2^lk: block size of this merge
        const uint32_t k = 1u << lk, hk = k >> 1;
        for (uint32_t t = threadIdx.x; t < half; t += nthreads) {
            const uint32_t blk = t >> (lk - 1), off = t & (hk - 1);
            const uint32_t i = (blk << lk) + off, l = (blk << lk) + (k - 1 - off);
            if (l < n) cmpxchg(a, i, l);
        }
        __syncthreads();
        for (int lj = (int)lk - 2; lj >= 0; --lj) {  // j = 2^lj: half-cleaner distance
            const uint32_t j = 1u << lj;
            for (uint32_t t = threadIdx.x; t < half; t += nthreads) {
                const uint32_t i = ((t >> lj) << (lj + 1)) + (t & (j - 1)), l = i + j;
                if (l < n) cmpxchg(a, i, l);
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(1024)
sort_tiles_kernel(int T, const uint32_t* __restrict__ starts, unsigned long long* __restrict__ keys,
                  uint32_t* __restrict__ points, uint32_t smem_cap, uint32_t capacity) {
    extern __shared__ __align__(16) unsigned long long sk[];
    pdl_enter();
    const int tile = blockIdx.x;
    if (tile >= T) return;
    const uint32_t s = min(starts[tile], capacity), e = min(starts[tile + 1], capacity), n = e - s;
    if (n == 0) return;
    const uint32_t nt = blockDim.x;  // 256, or 1024 when some tile exceeds the shared-memory tier
    unsigned long long* seg = keys + s;
    if (n <= smem_cap) {
        for (uint32_t i = threadIdx.x; i < n; i += nt) sk[i] = seg[i];
        __syncthreads();
        bitonic_sort(sk, n, nt);
        for (uint32_t i = threadIdx.x; i < n; i += nt) {
            const unsigned long long k = sk[i];
            seg[i] = k;
            points[s + i] = (uint32_t)k;
        }
    } else {  // oversized tile: same network directly on the global segment (block-scope visibility via the barriers)
        __syncthreads();
        bitonic_sort(seg, n, nt);
        for (uint32_t i = threadIdx.x; i < n; i += nt) points[s + i] = (uint32_t)seg[i];
    }
}

// ---------------------------------------------------------------------------------------
// Register-resident variant for the common case (every tile <= 256*WARPS pairs): each thread
// holds 8 consecutive keys, so compare-exchange distances 1, 2, 4 are thread-local, distances
// 8..128 are warp shuffles, and only distances >= 256 go through shared memory + a barrier
// (one step for 512 keys).  Same ascending-only "flip" network, same result.
// ---------------------------------------------------------------------------------------
constexpr int SORT_E = 8;  // keys per thread

__device__ __forceinline__ void cx(unsigned long long& a, unsigned long long& b) {  // ascending compare-exchange
    // one 64-bit compare, four selects (written out: the compiler otherwise emits a second compare for the maximum)
    unsigned long long lo, hi;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.u64 p, %2, %3;\n\t"
        "selp.b64 %0, %2, %3, p;\n\t"
        "selp.b64 %1, %3, %2, p;\n\t}"
        : "=l"(lo), "=l"(hi)
        : "l"(a), "l"(b));
    a = lo, b = hi;
}
__device__ __forceinline__ void local_tail(unsigned long long (&v)[SORT_E]) {  // distances 4, 2, 1
    cx(v[0], v[4]), cx(v[1], v[5]), cx(v[2], v[6]), cx(v[3], v[7]);
    cx(v[0], v[2]), cx(v[1], v[3]), cx(v[4], v[6]), cx(v[5], v[7]);
    cx(v[0], v[1]), cx(v[2], v[3]), cx(v[4], v[5]), cx(v[6], v[7]);
}
// keep the minimum (lower index side) or maximum of (mine, other)
__device__ __forceinline__ void keep(unsigned long long& mine, unsigned long long other, bool lower) {
    if ((other < mine) == lower) mine = other;
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
sort_tiles_reg_kernel(int T, const uint32_t* __restrict__ starts, unsigned long long* __restrict__ keys,
                      uint32_t* __restrict__ points, uint32_t capacity) {
    __shared__ unsigned long long xch[WARPS > 1 ? 256 * WARPS : 1];
    pdl_enter();
    const int tile = blockIdx.x;
    if (tile >= T) return;
    const uint32_t s = min(starts[tile], capacity), n = min(starts[tile + 1], capacity) - s;
    if (n == 0) return;
    unsigned long long* seg = keys + s;
    if (n > 256u * WARPS) {  // larger than the (hinted) register capacity: same network in place in global memory
        bitonic_sort(seg, n, 32 * WARPS);
        for (uint32_t i = threadIdx.x; i < n; i += 32 * WARPS) points[s + i] = (uint32_t)seg[i];
        return;
    }
    const uint32_t tid = threadIdx.x, lane = tid & 31, i0 = tid * SORT_E;
    int L = 3;  // log2 of the padded size (>= 8)
    while ((1u << L) < n) ++L;
    if (WARPS > 1 && L <= 8 && tid >= 32) return;  // one warp is enough and no barrier will be executed

    unsigned long long v[SORT_E];
#pragma unroll
    for (int e = 0; e < SORT_E; ++e) v[e] = (i0 + e < n) ? seg[i0 + e] : ~0ull;

    // levels 1..3 (blocks of 2, 4, 8 keys) are entirely thread-local
    cx(v[0], v[1]), cx(v[2], v[3]), cx(v[4], v[5]), cx(v[6], v[7]);
    cx(v[0], v[3]), cx(v[1], v[2]), cx(v[4], v[7]), cx(v[5], v[6]);
    cx(v[0], v[1]), cx(v[2], v[3]), cx(v[4], v[5]), cx(v[6], v[7]);
    cx(v[0], v[7]), cx(v[1], v[6]), cx(v[2], v[5]), cx(v[3], v[4]);
    cx(v[0], v[2]), cx(v[1], v[3]), cx(v[4], v[6]), cx(v[5], v[7]);
    cx(v[0], v[1]), cx(v[2], v[3]), cx(v[4], v[5]), cx(v[6], v[7]);

    for (int lk = 4; lk <= L; ++lk) {
        // flip step: key i pairs with i ^ (2^lk - 1): local slot e <-> 7-e in thread tid ^ (2^(lk-3) - 1)
        {
            const bool lower = ((tid >> (lk - 4)) & 1u) == 0;  // bit lk-1 of the key index
            unsigned long long o[SORT_E];
            if (WARPS == 1 || lk <= 8) {
                const int m = (1 << (lk - 3)) - 1;
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) o[e] = __shfl_xor_sync(0xffffffffu, v[SORT_E - 1 - e], m);
            } else {
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) xch[i0 + e] = v[e];
                __syncthreads();
                const uint32_t m = (1u << lk) - 1u;
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) o[e] = xch[(i0 + e) ^ m];
                __syncthreads();
            }
#pragma unroll
            for (int e = 0; e < SORT_E; ++e) keep(v[e], o[e], lower);
        }
        // half-cleaners with distance 2^lj, lj = lk-2 .. 3 (cross-thread), then the local tail (4, 2, 1)
        for (int lj = lk - 2; lj >= 3; --lj) {
            const bool lower = ((tid >> (lj - 3)) & 1u) == 0;  // bit lj of the key index
            unsigned long long o[SORT_E];
            if (WARPS == 1 || lj < 8) {
                const int m = 1 << (lj - 3);
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) o[e] = __shfl_xor_sync(0xffffffffu, v[e], m);
            } else {
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) xch[i0 + e] = v[e];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) o[e] = xch[(i0 + e) ^ (1u << lj)];
                __syncthreads();
            }
#pragma unroll
            for (int e = 0; e < SORT_E; ++e) keep(v[e], o[e], lower);
        }
        local_tail(v);
    }
#pragma unroll
    for (int e = 0; e < SORT_E; ++e)
        if (i0 + e < n) {
            seg[i0 + e] = v[e];
            points[s + i0 + e] = (uint32_t)v[e];
        }
    (void)lane;
}


// ---------------------------------------------------------------------------------------
// Quantised variant of the register sort (default): the network runs on 32-bit composites
//     (depth_bits - tile_min) >> shift  |  position of the key in the emitted segment
// instead of the 64-bit keys, so a compare-exchange is ONE shuffle + ONE VIMNMX (min or max chosen by a predicate)
// instead of two shuffles, a 64-bit compare and two selects -- a third of the instructions.  View depths are positive
// floats, so their bit patterns are monotone integers and the shift-quantisation is exactly monotone; the tile's
// own [min, max] range is spread over 32 - log2(capacity) bits (21 for tiles up to 2048 pairs).  Entries whose
// quantised depths collide (about 2 % of the tiles have one such pair at C2) come out adjacent and are ranked
// inside their run with the exact 64-bit keys (re-read from the unsorted segment through the position the composite
// carries), so the result is still the reference order, bit for bit.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cx32(uint32_t& a, uint32_t& b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo, b = hi;
}
__device__ __forceinline__ void local_tail32(uint32_t (&v)[SORT_E]) {  // distances 4, 2, 1
    cx32(v[0], v[4]), cx32(v[1], v[5]), cx32(v[2], v[6]), cx32(v[3], v[7]);
    cx32(v[0], v[2]), cx32(v[1], v[3]), cx32(v[4], v[6]), cx32(v[5], v[7]);
    cx32(v[0], v[1]), cx32(v[2], v[3]), cx32(v[4], v[5]), cx32(v[6], v[7]);
}
__device__ __forceinline__ uint32_t keep32(uint32_t mine, uint32_t other, bool lower) {
    return lower ? min(mine, other) : max(mine, other);
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
sort_tiles_q_kernel(int T, const uint32_t* __restrict__ starts, unsigned long long* __restrict__ keys,
                    uint32_t* __restrict__ points, uint32_t capacity) {
    constexpr int CAP = 256 * WARPS;
    constexpr int IDXB = WARPS == 1 ? 8 : WARPS == 2 ? 9 : WARPS == 4 ? 10 : WARPS == 8 ? 11 : WARPS == 16 ? 12 : 13;  // log2(CAP)
    constexpr uint32_t IDXM = (1u << IDXB) - 1u;
    __shared__ uint32_t xch[CAP];             // cross-warp exchange; afterwards the sorted composites
    __shared__ uint32_t smin[WARPS], smax[WARPS];
    pdl_enter();
    const int tile = blockIdx.x;
    if (tile >= T) return;
    const uint32_t s = min(starts[tile], capacity), n = min(starts[tile + 1], capacity) - s;
    if (n == 0) return;
    unsigned long long* seg = keys + s;
    if (n > (uint32_t)CAP) {  // larger than the (hinted) register capacity: the 64-bit network in place in global memory
        bitonic_sort(seg, n, 32 * WARPS);
        for (uint32_t i = threadIdx.x; i < n; i += 32 * WARPS) points[s + i] = (uint32_t)seg[i];
        return;
    }
    const uint32_t tid = threadIdx.x, warp = tid >> 5, i0 = tid * SORT_E;
    int L = 3;  // log2 of the padded size (>= 8)
    while ((1u << L) < n) ++L;
    const bool active = warp * 256u < (1u << L);  // warps beyond the padded size only keep the barriers company

    uint32_t d[SORT_E], dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
    for (int e = 0; e < SORT_E; ++e) {
        d[e] = 0u;
        if (i0 + e < n) {
            d[e] = (uint32_t)(seg[i0 + e] >> 32);
            dmin = min(dmin, d[e]), dmax = max(dmax, d[e]);
        }
    }
    dmin = __reduce_min_sync(0xffffffffu, dmin), dmax = __reduce_max_sync(0xffffffffu, dmax);
    if (WARPS > 1) {
        if ((tid & 31u) == 0) smin[warp] = dmin, smax[warp] = dmax;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < WARPS; ++w) dmin = min(dmin, smin[w]), dmax = max(dmax, smax[w]);
    }
    const int range_bits = 32 - __clz(dmax - dmin);  // 0 when all depths are equal
    const int shift = max(0, range_bits - (32 - IDXB));
    uint32_t v[SORT_E];
#pragma unroll
    for (int e = 0; e < SORT_E; ++e)
        v[e] = (i0 + e < n) ? ((((d[e] - dmin) >> shift) << IDXB) | (i0 + e)) : 0xffffffffu;

    if (active) {  // levels 1..3 (blocks of 2, 4, 8 keys) are entirely thread-local
        cx32(v[0], v[1]), cx32(v[2], v[3]), cx32(v[4], v[5]), cx32(v[6], v[7]);
        cx32(v[0], v[3]), cx32(v[1], v[2]), cx32(v[4], v[7]), cx32(v[5], v[6]);
        cx32(v[0], v[1]), cx32(v[2], v[3]), cx32(v[4], v[5]), cx32(v[6], v[7]);
        cx32(v[0], v[7]), cx32(v[1], v[6]), cx32(v[2], v[5]), cx32(v[3], v[4]);
        cx32(v[0], v[2]), cx32(v[1], v[3]), cx32(v[4], v[6]), cx32(v[5], v[7]);
        cx32(v[0], v[1]), cx32(v[2], v[3]), cx32(v[4], v[5]), cx32(v[6], v[7]);
    }
    for (int lk = 4; lk <= L; ++lk) {
        {  // flip step: key i pairs with i ^ (2^lk - 1): local slot e <-> 7-e in thread tid ^ (2^(lk-3) - 1)
            const bool lower = ((tid >> (lk - 4)) & 1u) == 0;  // bit lk-1 of the key index
            uint32_t o[SORT_E];
            if (WARPS == 1 || lk <= 8) {
                if (active) {
                    const int m = (1 << (lk - 3)) - 1;
#pragma unroll
                    for (int e = 0; e < SORT_E; ++e) o[e] = __shfl_xor_sync(0xffffffffu, v[SORT_E - 1 - e], m);
#pragma unroll
                    for (int e = 0; e < SORT_E; ++e) v[e] = keep32(v[e], o[e], lower);
                }
            } else {
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) xch[i0 + e] = v[e];
                __syncthreads();
                const uint32_t m = (1u << lk) - 1u;
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) o[e] = xch[(i0 + e) ^ m];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) v[e] = keep32(v[e], o[e], lower);
            }
        }
        // half-cleaners with distance 2^lj, lj = lk-2 .. 3 (cross-thread), then the local tail (4, 2, 1)
        for (int lj = lk - 2; lj >= 3; --lj) {
            const bool lower = ((tid >> (lj - 3)) & 1u) == 0;  // bit lj of the key index
            uint32_t o[SORT_E];
            if (WARPS == 1 || lj < 8) {
                if (active) {
                    const int m = 1 << (lj - 3);
#pragma unroll
                    for (int e = 0; e < SORT_E; ++e) o[e] = __shfl_xor_sync(0xffffffffu, v[e], m);
#pragma unroll
                    for (int e = 0; e < SORT_E; ++e) v[e] = keep32(v[e], o[e], lower);
                }
            } else {
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) xch[i0 + e] = v[e];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) o[e] = xch[(i0 + e) ^ (1u << lj)];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < SORT_E; ++e) v[e] = keep32(v[e], o[e], lower);
            }
        }
        if (active) local_tail32(v);
    }
    // sorted composites -> shared memory; the exact keys are read back from the (still unsorted) segment through the
    // position each composite carries, and entries whose quantised depth collides with a neighbour's are ranked inside
    // their run with those exact keys.  All reads of the segment happen before the barrier, all writes after it.
#pragma unroll
    for (int e = 0; e < SORT_E; ++e) xch[i0 + e] = v[e];
    if (WARPS > 1) __syncthreads(); else __syncwarp();
    unsigned long long key[SORT_E];
    uint32_t pos[SORT_E];
#pragma unroll
    for (int e = 0; e < SORT_E; ++e) {
        const uint32_t p = i0 + e;
        pos[e] = p, key[e] = 0ull;
        if (p >= n) continue;
        const uint32_t qd = v[e] >> IDXB;
        key[e] = seg[v[e] & IDXM];
        const bool tie_l = p > 0 && (xch[p - 1] >> IDXB) == qd, tie_r = p + 1 < n && (xch[p + 1] >> IDXB) == qd;
        if (tie_l || tie_r) {
            uint32_t a = p, b = p + 1;
            while (a > 0 && (xch[a - 1] >> IDXB) == qd) --a;
            while (b < n && (xch[b] >> IDXB) == qd) ++b;
            uint32_t rank = 0;
            for (uint32_t r = a; r < b; ++r) rank += seg[xch[r] & IDXM] < key[e] ? 1u : 0u;
            pos[e] = a + rank;
        }
    }
    if (WARPS > 1) __syncthreads(); else __syncwarp();
#pragma unroll
    for (int e = 0; e < SORT_E; ++e)
        if (i0 + e < n) {
            seg[pos[e]] = key[e];
            points[s + pos[e]] = (uint32_t)key[e];
        }
}

void launch_emit(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, uint32_t capacity, cudaStream_t s) {
    if (v.P == 0) return;
    launch_chain(emit_kernel, dim3((v.P + EMIT_THREADS - 1) / EMIT_THREADS), dim3(EMIT_THREADS), 0, s, v, g, im.cursor, b.keys, capacity);
}

void launch_sort_tiles(const View& v, ImagePtrs im, BinPtrs b, uint32_t max_tile_pairs, uint32_t capacity,
                       cudaStream_t s) {
    const int T = v.gx * v.gy;
#if GGRT_SORT_Q
    if (max_tile_pairs <= 256) return launch_chain(sort_tiles_q_kernel<1>, dim3(T), dim3(32), 0, s, T, im.starts, b.keys, b.points, capacity);
    if (max_tile_pairs <= 512) return launch_chain(sort_tiles_q_kernel<2>, dim3(T), dim3(64), 0, s, T, im.starts, b.keys, b.points, capacity);
    if (max_tile_pairs <= 1024) return launch_chain(sort_tiles_q_kernel<4>, dim3(T), dim3(128), 0, s, T, im.starts, b.keys, b.points, capacity);
    if (max_tile_pairs <= 2048) return launch_chain(sort_tiles_q_kernel<8>, dim3(T), dim3(256), 0, s, T, im.starts, b.keys, b.points, capacity);
    if (max_tile_pairs <= 4096) return launch_chain(sort_tiles_q_kernel<16>, dim3(T), dim3(512), 0, s, T, im.starts, b.keys, b.points, capacity);
    if (max_tile_pairs <= 8192) return launch_chain(sort_tiles_q_kernel<32>, dim3(T), dim3(1024), 0, s, T, im.starts, b.keys, b.points, capacity);
#endif
    if (max_tile_pairs <= 256) {
        launch_chain(sort_tiles_reg_kernel<1>, dim3(T), dim3(32), 0, s, T, im.starts, b.keys, b.points, capacity);
        return;
    }
    if (max_tile_pairs <= 512) {
        launch_chain(sort_tiles_reg_kernel<2>, dim3(T), dim3(64), 0, s, T, im.starts, b.keys, b.points, capacity);
        return;
    }
    if (max_tile_pairs <= 1024) {
        launch_chain(sort_tiles_reg_kernel<4>, dim3(T), dim3(128), 0, s, T, im.starts, b.keys, b.points, capacity);
        return;
    }
    if (max_tile_pairs <= 2048) {
        launch_chain(sort_tiles_reg_kernel<8>, dim3(T), dim3(256), 0, s, T, im.starts, b.keys, b.points, capacity);
        return;
    }
    // shared-memory capacity tier from the largest tile (reported by scan_tiles)
    uint32_t cap = 1024;
    while (cap < max_tile_pairs && cap < 16384) cap <<= 1;
    const size_t smem = (size_t)cap * sizeof(unsigned long long);
    if (smem > 32 * 1024)  // per-device attribute; cheap host-side call
        cudaFuncSetAttribute(sort_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    // a tile beyond the shared-memory tier is sorted in place in global memory by its CTA: give it the largest CTA
    const int threads = max_tile_pairs > 16384 ? 1024 : SORT_THREADS;
    launch_chain(sort_tiles_kernel, dim3(v.gx * v.gy), dim3(threads), smem, s, v.gx * v.gy, im.starts, b.keys, b.points, cap, capacity);
}

}  // namespace ggrt
