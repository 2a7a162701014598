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

constexpr int EMIT_THREADS = 256;
constexpr int SORT_THREADS = 256;

__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(View v, GeomPtrs g, uint32_t* __restrict__ cursor, unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * EMIT_THREADS + threadIdx.x;
    if (i >= v.P) return;
    if (g.tiles[i] == 0) return;
    const ushort4 r = g.rect[i];
    const unsigned long long key = ((unsigned long long)__float_as_uint(g.rec2[i].w) << 32) | (uint32_t)i;
    const int sub = i & (SUBS - 1);
    for (int y = r.y; y < r.w; ++y)
        for (int x = r.x; x < r.z; ++x) {
            // the (tile, sub-counter) segment start doubles as its allocation cursor: one returning atomic per pair
            const uint32_t slot = atomicAdd(&cursor[((y * v.gx + x) << SUBS_LOG2) + sub], 1u);
            keys[slot] = key;
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
__device__ __forceinline__ void bitonic_sort(unsigned long long* a, uint32_t n) {
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    const uint32_t half = n2 >> 1;
    for (uint32_t lk = 1; (1u << lk) <= n2; ++lk) {  // k = 2^lk: block size of this merge
        const uint32_t k = 1u << lk, hk = k >> 1;
        for (uint32_t t = threadIdx.x; t < half; t += SORT_THREADS) {
            const uint32_t blk = t >> (lk - 1), off = t & (hk - 1);
            const uint32_t i = (blk << lk) + off, l = (blk << lk) + (k - 1 - off);
            if (l < n) cmpxchg(a, i, l);
        }
        __syncthreads();
        for (int lj = (int)lk - 2; lj >= 0; --lj) {  // j = 2^lj: half-cleaner distance
            const uint32_t j = 1u << lj;
            for (uint32_t t = threadIdx.x; t < half; t += SORT_THREADS) {
                const uint32_t i = ((t >> lj) << (lj + 1)) + (t & (j - 1)), l = i + j;
                if (l < n) cmpxchg(a, i, l);
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_tiles_kernel(int T, const uint32_t* __restrict__ starts, unsigned long long* __restrict__ keys,
                  uint32_t* __restrict__ points, uint32_t smem_cap) {
    extern __shared__ __align__(16) unsigned long long sk[];
    const int tile = blockIdx.x;
    if (tile >= T) return;
    const uint32_t s = starts[tile], e = starts[tile + 1], n = e - s;
    if (n == 0) return;
    unsigned long long* seg = keys + s;
    if (n <= smem_cap) {
        for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) sk[i] = seg[i];
        __syncthreads();
        bitonic_sort(sk, n);
        for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) {
            const unsigned long long k = sk[i];
            seg[i] = k;
            points[s + i] = (uint32_t)k;
        }
    } else {  // oversized tile: same network directly on the global segment (block-scope visibility via the barriers)
        __syncthreads();
        bitonic_sort(seg, n);
        for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) points[s + i] = (uint32_t)seg[i];
    }
}

void launch_emit(const View& v, const int*, GeomPtrs g, ImagePtrs im, BinPtrs b, cudaStream_t s) {
    if (v.P == 0) return;
    emit_kernel<<<(v.P + EMIT_THREADS - 1) / EMIT_THREADS, EMIT_THREADS, 0, s>>>(v, g, im.cursor, b.keys);
}

void launch_sort_tiles(const View& v, ImagePtrs im, BinPtrs b, uint32_t max_tile_pairs, cudaStream_t s) {
    // shared-memory capacity tier from the largest tile (reported by scan_tiles)
    uint32_t cap = 1024;
    while (cap < max_tile_pairs && cap < 16384) cap <<= 1;
    const size_t smem = (size_t)cap * sizeof(unsigned long long);
    if (smem > 32 * 1024)  // per-device attribute; cheap host-side call
        cudaFuncSetAttribute(sort_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    sort_tiles_kernel<<<v.gx * v.gy, SORT_THREADS, smem, s>>>(v.gx * v.gy, im.starts, b.keys, b.points, cap);
}

}  // namespace ggrt
