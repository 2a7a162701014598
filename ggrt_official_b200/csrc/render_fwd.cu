// Per-tile front-to-back alpha compositing (A.3).  Replaces upstream renderCUDA (fwd),
// SURVEY.md 8a row a11.
//
// One CTA per 16x16 tile; warp w owns the 8x4 pixel block (w&1, w>>1) so each warp's
// output rows are 32-byte contiguous.  A batch of 256 Gaussian records (48 B each) is
// staged into shared memory; the thread that staged a record culls it against all 8 warp
// pixel blocks of the tile (exact ellipse-vs-rectangle test on {alpha >= 1/255}, block_mask8)
// and publishes the 8-bit result in shared memory and, for the backward kernel, in the pair
// buffer; each warp then ballots its bit and walks only the surviving entries, in list order.
// GGRt's splats are a few pixels wide, so most (warp, Gaussian) pairs of a tile are culled.
#include "f32x2.cuh"
#include "render_common.cuh"
#include "tma.cuh"

namespace ggrt {

// ASYNC (north_star: "TMA/cp.async staging of per-tile Gaussian records into shared memory"): the records of batch
// b+1 are gathered with cp.async (SASS: LDGSTS, 3 x 16 B per record, no register round trip) into the other half
// of a double buffer while batch b is culled and blended; the list index a thread needs for that gather is itself
// loaded one batch ahead.  The conic is rescaled for the ex2 exponent by a short in-place pass once a batch has
// landed.  Measured against the synchronous fill in DESIGN.md section 8.
#ifndef GGRT_FWD_ASYNC
#define GGRT_FWD_ASYNC 1
#endif
// GGRT_FWD_PIPE=1: per-warp survivor queue + software-pipelined walk (entry i+1 evaluated while entry i is blended).
// Measured on B200 at C2: 101 us against 81 us for the ballot walk below (64 instead of 46 registers, one wasted
// evaluation per batch and warp, coarser early termination) -- kept for A/B builds only.
#ifndef GGRT_FWD_PIPE
#define GGRT_FWD_PIPE 0
#endif
#ifndef GGRT_FWD_MINBLOCKS
#define GGRT_FWD_MINBLOCKS 4
#endif
template <bool ASYNC>
__global__ void __launch_bounds__(FWD_THREADS, GGRT_FWD_MINBLOCKS * 256 / FWD_THREADS)
render_forward_kernel(View v, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                      const float4* __restrict__ rec2, const uint32_t* __restrict__ starts,
                      const uint32_t* __restrict__ points, uint8_t* __restrict__ masks, uint32_t capacity,
                      float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ final_T,
                      uint32_t* __restrict__ n_contrib) {
    __shared__ __align__(16) unsigned char srec[(ASYNC ? 2 : 1) * FWD_BATCH * REC_BYTES];
    __shared__ __align__(4) uint8_t smask[FWD_BATCH];  // per staged record: the warp pixel blocks it reaches
#if GGRT_FWD_PIPE
    __shared__ uint8_t squeue[FWD_WARPS][FWD_BATCH];   // per warp: the batch entries that reach its pixel block
#endif
    uint32_t sbase = smem_addr(srec);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int wt = warp + blockIdx.z * FWD_WARPS;  // warp pixel block of the tile (8 per tile)
    const int bx0 = blockIdx.x * TILE + (wt & 1) * 8, by0 = blockIdx.y * TILE + (wt >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = px < v.W && py < v.H;
    const float pxf = (float)px, pyf = (float)py;
    const float tx0f = (float)(blockIdx.x * TILE), ty0f = (float)(blockIdx.y * TILE);
    const uint32_t start = min(starts[tile], capacity), end = min(starts[tile + 1], capacity);

    // The sign of T carries the per-pixel "done" flag (T > 0: live, T < 0: terminated with final transmittance |T|;
    // a live T never drops below T_EPS), which removes the predicate bookkeeping from the blend loop: a terminated
    // pixel yields Tn < 0 and therefore never blends again.
    float T = inside ? 1.0f : -1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t last = 0;

    static_assert(FWD_THREADS == FWD_BATCH, "one record per thread and batch");
    // ASYNC: gather of batch `b` into half `h` of the double buffer; nid = this thread's list entry of the batch after
    uint32_t nid = 0;
    auto gather = [&](uint32_t b, uint32_t h) {
        if (b + tid < end) {
            const uint32_t dst = smem_addr(srec) + h * (FWD_BATCH * REC_BYTES) + tid * REC_BYTES;
            cp_async16(dst, rec0 + nid);
            cp_async16(dst + 16, rec1 + nid);
            cp_async16(dst + 32, rec2 + nid);
        }
        cp_async_commit();
        const uint32_t nn = b + FWD_BATCH + tid;
        if (nn < end) nid = points[nn];
    };
    if (ASYNC) {
        if (start + tid < end) nid = points[start + tid];
        gather(start, 0);
    }
    uint32_t half = 0;
    for (uint32_t base = start; base < end; base += FWD_BATCH, half ^= 1u) {
        if (__syncthreads_and(T < 0.0f)) break;  // also orders the previous batch's reads before the refill
        const uint32_t cnt = min((uint32_t)FWD_BATCH, end - base);
        if (ASYNC) {
            sbase = smem_addr(srec) + half * (FWD_BATCH * REC_BYTES);
            gather(base + FWD_BATCH, half ^ 1u);   // the next batch streams in under this one's blend loop
            cp_async_wait<1>();                    // this thread's copies of the current batch have landed ...
            if (tid < cnt) {  // ... cull it against the tile's 8 warp pixel blocks (ONE exact test per record and block for
                              // the whole CTA; the backward kernel reuses the result) and rescale its conic for the ex2
                              // exponent, in place
                const float4 a = lds128(sbase + tid * REC_BYTES);
                float4 c = lds128(sbase + tid * REC_BYTES + 16);
                const uint32_t m = block_mask8(a.x, a.y, a.z, c.x, c.y, c.z, tx0f, ty0f);
                smask[tid] = (uint8_t)m;
                masks[base + tid] = (uint8_t)m;
                c.x *= -0.5f * LOG2E, c.y *= -LOG2E, c.z *= -0.5f * LOG2E;
                sts128(sbase + tid * REC_BYTES + 16, c);
            }
        } else {
            for (uint32_t k = tid; k < cnt; k += FWD_THREADS) {
                const uint32_t id = points[base + k];
                const uint32_t dst = sbase + k * REC_BYTES;
                const float4 a = rec0[id];
                float4 c = rec1[id];
                const uint32_t m = block_mask8(a.x, a.y, a.z, c.x, c.y, c.z, tx0f, ty0f);
                smask[k] = (uint8_t)m;
                masks[base + k] = (uint8_t)m;
                c.x *= -0.5f * LOG2E, c.y *= -LOG2E, c.z *= -0.5f * LOG2E;
                sts128(dst, a);
                sts128(dst + 16, c);
                sts128(dst + 32, rec2[id]);
            }
        }
        __syncthreads();
        if (__all_sync(0xffffffffu, T < 0.0f)) continue;
#if GGRT_FWD_PIPE
        // ---- the entries that reach this warp's pixel block, compacted in list order ----------------------------
        uint32_t qn = 0;
        for (uint32_t r = 0; r < cnt; r += 32) {
            const uint32_t j = r + lane;
            const bool hit = j < cnt && ((smask[j] >> wt) & 1u);
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) squeue[warp][qn + __popc(m & ((1u << lane) - 1u))] = (uint8_t)j;
            qn += __popc(m);
        }
        __syncwarp();
        // ---- software-pipelined walk: the Gaussian of entry i+1 is evaluated (record loads, exponent, ex2, alpha) while
        // entry i is blended -- the blend is a dependent chain through T, the evaluation is not, so the two interleave
        // and the shared-memory / MUFU latencies of the evaluation no longer stall the warp ------------------------
        if (qn) {
            const uint32_t pos0 = (base - start) + 1;  // 1-based list position of the batch's entry 0
            auto eval = [&](uint32_t j, float& al, float4& col) {
                const uint32_t src = sbase + j * REC_BYTES;
                const float2 xy = lds64(src);
                const float4 c = lds128(src + 16);
                col = lds128(src + 32);
                const float dx = xy.x - pxf, dy = xy.y - pyf;
                const float power2 = fmaf(dx, fmaf(c.x, dx, c.y * dy), (c.z * dy) * dy);
                const float alpha = fminf(ALPHA_MAX, c.w * ex2_approx(power2));
                al = ((power2 <= 0.0f) && (alpha >= ALPHA_MIN)) ? alpha : 0.0f;  // 0: not active at this pixel
            };
            uint32_t j = squeue[warp][0];
            float al;
            float4 col;
            eval(j, al, col);
            for (uint32_t i = 0; i < qn; ++i) {
                const uint32_t jn = squeue[warp][min(i + 1, qn - 1)];
                float aln;
                float4 coln;
                eval(jn, aln, coln);
                const bool active = al > 0.0f;
                const float Tn = T * (1.0f - al);
                const bool blend = active && (Tn >= T_EPS);
                const float w = blend ? al * T : 0.0f;
                C0 = fmaf(col.x, w, C0);
                C1 = fmaf(col.y, w, C1);
                C2 = fmaf(col.z, w, C2);
                D = fmaf(col.w, w, D);
                const float Tstop = active ? -fabsf(T) : T;  // a live pixel that cannot blend an active Gaussian stops
                T = blend ? Tn : Tstop;
                last = blend ? pos0 + j : last;
                al = aln, col = coln, j = jn;
                if ((i & 7u) == 7u && __all_sync(0xffffffffu, T < 0.0f)) break;
            }
        }
#else
        for (uint32_t r = 0; r < cnt; r += 32) {
            // lane l tests list entry r + 31 - l: bit b of the ballot is entry r + 31 - b, the highest bit comes first
            const uint32_t j = r + 31 - lane;
            const bool hit = j < cnt && ((smask[j] >> wt) & 1u);
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            const uint32_t top = sbase + (r + 31) * REC_BYTES;       // record of bit 0 ... minus b records for bit b
            const uint32_t last_top = (base - start) + r + 32;       // 1-based list position of bit 0's entry ... - b
            while (mask) {
                uint32_t b, src;  // opaque to the optimiser, which otherwise rebuilds the index from 31 - clz
                asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(mask));
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(src) : "r"(b), "r"(0u - (uint32_t)REC_BYTES), "r"(top));
                mask ^= 1u << b;
                const float2 xy = lds64(src);
                const float4 c = lds128(src + 16);
                const float4 col = lds128(src + 32);
                const float dx = xy.x - pxf, dy = xy.y - pyf;
                const float power2 = fmaf(dx, fmaf(c.x, dx, c.y * dy), (c.z * dy) * dy);
                const float alpha = fminf(ALPHA_MAX, c.w * ex2_approx(power2));
                const bool active = (power2 <= 0.0f) && (alpha >= ALPHA_MIN);
                const float Tn = T * (1.0f - alpha);
                const bool blend = active && (Tn >= T_EPS);
                const float w = blend ? alpha * T : 0.0f;
                C0 = fmaf(col.x, w, C0);
                C1 = fmaf(col.y, w, C1);
                C2 = fmaf(col.z, w, C2);
                D = fmaf(col.w, w, D);
                const float Tstop = active ? -fabsf(T) : T;  // a live pixel that cannot blend an active Gaussian stops
                T = blend ? Tn : Tstop;
                last = blend ? last_top - b : last;
            }
            if (__all_sync(0xffffffffu, T < 0.0f)) break;
        }
#endif
    }
    if (ASYNC) cp_async_wait<0>();  // nothing may still be in flight into this CTA's shared memory when it exits
    if (inside) {
        const float Tf = fabsf(T);
        const size_t pix = (size_t)py * v.W + px, hw = (size_t)v.H * v.W;
        out_color[pix] = fmaf(Tf, v.bg[0], C0);
        out_color[hw + pix] = fmaf(Tf, v.bg[1], C1);
        out_color[2 * hw + pix] = fmaf(Tf, v.bg[2], C2);
        out_depth[pix] = D;
        final_T[pix] = Tf;
        n_contrib[pix] = last;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Pixel-pair variant (GGRT_FWD_PAIR, default): one CTA of 4 warps per tile, warp w owns the 8x8 pixel block
// (w & 1, w >> 1) and every lane TWO pixels of it, (x, y) and (x, y + 4), which share dx and whose dy differ by the
// constant 4 -- so the blend body runs on packed f32x2 pairs (FFMA2 / FMUL2: one issue slot, two pixels).  The kernel
// is instruction-issue bound and the body is 75 % of its instructions: per list entry 1.21 8x8 blocks survive the cull
// and cost ~45 instructions each, against 1.78 8x4 blocks at 34 (CPU model tools/model/cull_shapes.py).  Each packed
// operation performs exactly the two scalar operations of the single-pixel kernel, so the image is bit-identical.
// Batches of 128 records (one per thread), cp.async double buffer and the 8-bit cull masks for the backward kernel as
// above; a warp's own cull bit is the OR of the two 8x4 blocks it covers.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef GGRT_FWD_PAIR_RPT
#define GGRT_FWD_PAIR_RPT 1
#endif
constexpr int F2_WARPS = 4, F2_THREADS = 32 * F2_WARPS;
constexpr int F2_RPT = GGRT_FWD_PAIR_RPT;  // records staged per thread and batch
constexpr int F2_BATCH = F2_THREADS * F2_RPT;
#ifndef GGRT_FWD_PAIR
#define GGRT_FWD_PAIR 1
#endif
#ifndef GGRT_FWD_PAIR_MINBLOCKS
#define GGRT_FWD_PAIR_MINBLOCKS 8
#endif

__global__ void __launch_bounds__(F2_THREADS, GGRT_FWD_PAIR_MINBLOCKS)
render_forward_pair_kernel(View v, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                           const float4* __restrict__ rec2, const uint32_t* __restrict__ starts,
                           const uint32_t* __restrict__ points, uint8_t* __restrict__ masks, uint32_t capacity,
                           float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ final_T,
                           uint32_t* __restrict__ n_contrib) {
    __shared__ __align__(16) unsigned char srec[2 * F2_BATCH * REC_BYTES];
    __shared__ __align__(4) uint8_t smask[F2_BATCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int bx0 = blockIdx.x * TILE + (warp & 1) * 8, by0 = blockIdx.y * TILE + (warp >> 1) * 8;
    const int px = bx0 + (lane & 7), pya = by0 + (lane >> 3), pyb = pya + 4;
    const bool ina = px < v.W && pya < v.H, inb = px < v.W && pyb < v.H;
    const float pxf = (float)px, pyaf = (float)pya;
    const float tx0f = (float)(blockIdx.x * TILE), ty0f = (float)(blockIdx.y * TILE);
    // the two 8x4 blocks of the cull mask that make up this warp's 8x8 block: rows 2 (w >> 1) and 2 (w >> 1) + 1
    const uint32_t wbits = (1u << (4 * (warp >> 1) + (warp & 1))) | (1u << (4 * (warp >> 1) + 2 + (warp & 1)));
    pdl_enter();  // (everything above is index arithmetic: it runs while the sort kernel drains)
    const uint32_t start = min(starts[tile], capacity), end = min(starts[tile + 1], capacity);

    // sign of T = per-pixel "done" flag, as in the single-pixel kernel
    f2 T2 = pk(ina ? 1.0f : -1.0f, inb ? 1.0f : -1.0f), C0 = bc(0.f), C1 = bc(0.f), C2 = bc(0.f), D = bc(0.f);
    uint32_t lasta = 0, lastb = 0;

    uint32_t nid[F2_RPT];
    auto gather = [&](uint32_t b, uint32_t h) {
#pragma unroll
        for (int u = 0; u < F2_RPT; ++u) {
            const uint32_t k = tid + u * F2_THREADS;
            if (b + k < end) {
                const uint32_t dst = smem_addr(srec) + h * (F2_BATCH * REC_BYTES) + k * REC_BYTES;
                cp_async16(dst, rec0 + nid[u]);
                cp_async16(dst + 16, rec1 + nid[u]);
                cp_async16(dst + 32, rec2 + nid[u]);
            }
        }
        cp_async_commit();
#pragma unroll
        for (int u = 0; u < F2_RPT; ++u) {
            const uint32_t nn = b + F2_BATCH + tid + u * F2_THREADS;
            if (nn < end) nid[u] = points[nn];
        }
    };
#pragma unroll
    for (int u = 0; u < F2_RPT; ++u) {
        nid[u] = 0;
        if (start + tid + u * F2_THREADS < end) nid[u] = points[start + tid + u * F2_THREADS];
    }
    gather(start, 0);
    uint32_t half = 0;
    for (uint32_t base = start; base < end; base += F2_BATCH, half ^= 1u) {
        const bool done = lo(T2) < 0.0f && hi(T2) < 0.0f;
        if (__syncthreads_and(done)) break;  // also orders the previous batch's reads before the refill
        const uint32_t cnt = min((uint32_t)F2_BATCH, end - base);
        const uint32_t sbase = smem_addr(srec) + half * (F2_BATCH * REC_BYTES);
        gather(base + F2_BATCH, half ^ 1u);  // the next batch streams in under this one's blend loop
        cp_async_wait<1>();
#pragma unroll
        for (int u = 0; u < F2_RPT; ++u) {
            const uint32_t k = tid + u * F2_THREADS;
            if (k < cnt) {  // cull against the tile's 8 warp pixel blocks (8x4; the backward kernel's granularity), rescale
                const float4 a = lds128(sbase + k * REC_BYTES);
                float4 c = lds128(sbase + k * REC_BYTES + 16);
                const uint32_t m = block_mask8(a.x, a.y, a.z, c.x, c.y, c.z, tx0f, ty0f);
                smask[k] = (uint8_t)m;
                masks[base + k] = (uint8_t)m;
                c.x *= -0.5f * LOG2E, c.y *= -LOG2E, c.z *= -0.5f * LOG2E;
                sts128(sbase + k * REC_BYTES + 16, c);
            }
        }
        __syncthreads();
        if (__all_sync(0xffffffffu, lo(T2) < 0.0f && hi(T2) < 0.0f)) continue;
        for (uint32_t r = 0; r < cnt; r += 32) {
            // lane l tests list entry r + 31 - l: bit b of the ballot is entry r + 31 - b, the highest bit comes first
            const uint32_t j = r + 31 - lane;
            const bool hit = j < cnt && (smask[j] & wbits);
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            const uint32_t top = sbase + (r + 31) * REC_BYTES;       // record of bit 0 ... minus b records for bit b
            const uint32_t last_top = (base - start) + r + 32;       // 1-based list position of bit 0's entry ... - b
            while (mask) {
                uint32_t b, src;  // opaque to the optimiser, which otherwise rebuilds the index from 31 - clz
                asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(mask));
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(src) : "r"(b), "r"(0u - (uint32_t)REC_BYTES), "r"(top));
                mask ^= 1u << b;
                const float2 xy = lds64(src);
                const float4 c = lds128(src + 16);
                const float4 col = lds128(src + 32);
                const float dx = xy.x - pxf, dya = xy.y - pyaf;
                const f2 dy2 = pk(dya, dya - 4.0f);
                // power2 = fma(dx, fma(c.x, dx, c.y * dy), (c.z * dy) * dy), both pixels at once
                const f2 power2 = fma2(bc(dx), fma2(bc(c.x), bc(dx), mul2(bc(c.y), dy2)), mul2(mul2(bc(c.z), dy2), dy2));
                const float pa = lo(power2), pb = hi(power2);
                const f2 raw2 = mul2(bc(c.w), pk(ex2_approx(pa), ex2_approx(pb)));
                const float ala = fminf(ALPHA_MAX, lo(raw2)), alb = fminf(ALPHA_MAX, hi(raw2));
                const bool acta = (pa <= 0.0f) && (ala >= ALPHA_MIN), actb = (pb <= 0.0f) && (alb >= ALPHA_MIN);
                const f2 al2 = pk(ala, alb);
                const f2 Tn2 = mul2(T2, fma2(al2, bc(-1.0f), bc(1.0f)));  // T (1 - alpha)
                const float Ta = lo(T2), Tb = hi(T2), Tna = lo(Tn2), Tnb = hi(Tn2);
                const bool bla = acta && (Tna >= T_EPS), blb = actb && (Tnb >= T_EPS);
                const f2 aT2 = mul2(al2, T2);
                const f2 w2 = pk(bla ? lo(aT2) : 0.0f, blb ? hi(aT2) : 0.0f);
                C0 = fma2(bc(col.x), w2, C0);
                C1 = fma2(bc(col.y), w2, C1);
                C2 = fma2(bc(col.z), w2, C2);
                D = fma2(bc(col.w), w2, D);
                const float Tsa = acta ? -fabsf(Ta) : Ta, Tsb = actb ? -fabsf(Tb) : Tb;  // live pixel that cannot blend stops
                T2 = pk(bla ? Tna : Tsa, blb ? Tnb : Tsb);
                const uint32_t pos = last_top - b;
                lasta = bla ? pos : lasta;
                lastb = blb ? pos : lastb;
            }
            if (__all_sync(0xffffffffu, lo(T2) < 0.0f && hi(T2) < 0.0f)) break;
        }
    }
    cp_async_wait<0>();  // nothing may still be in flight into this CTA's shared memory when it exits
    const size_t hw = (size_t)v.H * v.W;
    if (ina) {
        const float Tf = fabsf(lo(T2));
        const size_t pix = (size_t)pya * v.W + px;
        out_color[pix] = fmaf(Tf, v.bg[0], lo(C0));
        out_color[hw + pix] = fmaf(Tf, v.bg[1], lo(C1));
        out_color[2 * hw + pix] = fmaf(Tf, v.bg[2], lo(C2));
        out_depth[pix] = lo(D);
        final_T[pix] = Tf;
        n_contrib[pix] = lasta;
    }
    if (inb) {
        const float Tf = fabsf(hi(T2));
        const size_t pix = (size_t)pyb * v.W + px;
        out_color[pix] = fmaf(Tf, v.bg[0], hi(C0));
        out_color[hw + pix] = fmaf(Tf, v.bg[1], hi(C1));
        out_color[2 * hw + pix] = fmaf(Tf, v.bg[2], hi(C2));
        out_depth[pix] = hi(D);
        final_T[pix] = Tf;
        n_contrib[pix] = lastb;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-specialised persistent variant (GGRT_FWD_WS): the pixel-pair kernel above spends ~23 % of its warp time in the
// per-tile prologue -- three dependent global round trips (tile range -> list entries -> records) and the cull-mask pass,
// with all four warps waiting.  Here every CTA is persistent (tiles blockIdx.x, + gridDim.x, ...), has ONE producer warp
// and four consumer warps, and a 3-stage shared-memory ring between them (mbarrier full / empty pairs):
//   producer: for each tile and batch of 128 list entries: wait for a free stage, load the entries, gather their records
//             with cp.async, cull them against the tile's 8 warp pixel blocks (block_mask8; also written to the pair
//             buffer for the backward kernel), rescale the conics, publish {tile, offset, count, last} and arrive on `full`;
//   consumers: wait on `full`, blend the batch exactly as the pixel-pair kernel does, arrive on `empty`; the image is
//             written when a tile's last batch has been consumed.
// The producer runs up to three batches -- usually two tiles -- ahead, so its latencies never stall a consumer.  When all
// pixels of a tile have terminated the consumers say so (done_tile) and the producer closes the tile with an empty batch.
// ---------------------------------------------------------------------------------------------------------------------
// Measured on B200 at C2 (round 2): 99 us against 77 us for the pixel-pair kernel (101 us with a producer that
// completed one batch before gathering the next) -- the consumers wait less, but there are only 24 of them per SM
// instead of 32 (64 registers x 160 threads), the four waiters of a stage spin on the mbarrier, and 3024 tiles over
// 888 persistent CTAs quantise to 4 tile times where 3.4 would be ideal.  Kept for A/B builds (-DGGRT_FWD_WS=1).
#ifndef GGRT_FWD_WS
#define GGRT_FWD_WS 0
#endif
#ifndef GGRT_FWD_WS_MINBLOCKS
#define GGRT_FWD_WS_MINBLOCKS 6
#endif
#ifndef GGRT_FWD_WS_STAGES
#define GGRT_FWD_WS_STAGES 3
#endif
constexpr int WS_CONS = 4, WS_THREADS = 32 * (WS_CONS + 1), WS_BATCH = 128, WS_STAGES = GGRT_FWD_WS_STAGES;
struct __align__(16) WsMeta {
    int tile;       // -1: no more work
    uint32_t off;   // list offset of the batch inside its tile
    uint32_t cnt;   // records staged (0: empty tile, or a tile closed early)
    int last;       // the tile ends with this batch
};

__device__ __forceinline__ bool consumers_all(bool pred) {  // barrier 1 over the four consumer warps, AND-reduction
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "barrier.red.and.pred p, 1, %2, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"((uint32_t)pred), "n"(32 * WS_CONS)
        : "memory");
    return r != 0;
}

__global__ void __launch_bounds__(WS_THREADS, GGRT_FWD_WS_MINBLOCKS)
render_forward_ws_kernel(View v, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                         const float4* __restrict__ rec2, const uint32_t* __restrict__ starts,
                         const uint32_t* __restrict__ points, uint8_t* __restrict__ masks, uint32_t capacity,
                         float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ final_T,
                         uint32_t* __restrict__ n_contrib) {
    __shared__ __align__(16) unsigned char srec[WS_STAGES * WS_BATCH * REC_BYTES];
    __shared__ __align__(4) uint8_t smask[WS_STAGES][WS_BATCH];
    __shared__ WsMeta smeta[WS_STAGES];
    __shared__ __align__(8) unsigned long long full_bar[WS_STAGES], empty_bar[WS_STAGES];
    __shared__ int done_tile_s;
    volatile int* done_tile = &done_tile_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = v.gx * v.gy;
    if (tid == 0) {
        for (int st = 0; st < WS_STAGES; ++st) mbar_init(smem_u32(&full_bar[st]), 1), mbar_init(smem_u32(&empty_bar[st]), WS_CONS);
        done_tile_s = -1;
    }
    __syncthreads();
    uint32_t stage = 0, phase = 0;

    if (warp == WS_CONS) {
        // ------------------------------------------------ producer ------------------------------------------------
        // Two batches in flight: the gather of item i+1 (list entries, then cp.async of the records) is issued before
        // the cull pass of item i, so the producer's own round trips overlap; the next tile's range is loaded a tile ahead.
        int tile = blockIdx.x;
        uint32_t start = 0, end = 0, base = 0, nstart = 0, nend = 0;
        auto load_next_range = [&](int t) {
            if (t < T) nstart = min(starts[t], capacity), nend = min(starts[t + 1], capacity);
        };
        if (tile < T) start = min(starts[tile], capacity), end = min(starts[tile + 1], capacity);
        base = start;
        load_next_range(tile + (int)gridDim.x);
        // an issued item: its ring stage, meta data, list position and tile origin
        uint32_t p_stage[2] = {0u, 0u}, p_gbase[2] = {0u, 0u};
        WsMeta p_meta[2];
        uint32_t istage = 0, iphase = 0;  // ring position of the next issue (stage / phase above: next publish)
        int np = 0, head = 0;
        bool more = tile < T, sentinel = false;
        for (;;) {
            while (np < 2 && (more || !sentinel)) {
                mbar_wait(smem_u32(&empty_bar[istage]), iphase ^ 1u);
                const int slot = (head + np) & 1;
                p_stage[slot] = istage;
                if (!more) {  // all tiles issued: the end marker
                    p_meta[slot] = WsMeta{-1, 0u, 0u, 1};
                    p_gbase[slot] = 0u;
                    sentinel = true;
                } else {
                    // consumers report a tile whose pixels have all terminated: close it with an empty batch
                    const bool stop = base > start && __shfl_sync(0xffffffffu, *done_tile == tile ? 1 : 0, 0);
                    const uint32_t cnt = (stop || base >= end) ? 0u : min((uint32_t)WS_BATCH, end - base);
                    const bool last = stop || base + WS_BATCH >= end;
                    p_meta[slot] = WsMeta{tile, base - start, cnt, last ? 1 : 0};
                    p_gbase[slot] = base;
                    const uint32_t sb = smem_addr(srec) + istage * (WS_BATCH * REC_BYTES);
                    uint32_t id[WS_BATCH / 32];
#pragma unroll
                    for (int u = 0; u < WS_BATCH / 32; ++u) {
                        const uint32_t k = lane + 32u * u;
                        id[u] = k < cnt ? points[base + k] : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < WS_BATCH / 32; ++u) {
                        const uint32_t k = lane + 32u * u;
                        if (k < cnt) {
                            cp_async16(sb + k * REC_BYTES, rec0 + id[u]);
                            cp_async16(sb + k * REC_BYTES + 16, rec1 + id[u]);
                            cp_async16(sb + k * REC_BYTES + 32, rec2 + id[u]);
                        }
                    }
                    if (last) {
                        tile += (int)gridDim.x;
                        start = nstart, end = nend, base = start;
                        more = tile < T;
                        load_next_range(tile + (int)gridDim.x);
                    } else {
                        base += WS_BATCH;
                    }
                }
                cp_async_commit();  // one group per issued item (possibly empty)
                if (++istage == (uint32_t)WS_STAGES) istage = 0, iphase ^= 1u;
                ++np;
            }
            if (np == 0) break;
            // complete the oldest item in flight: its records have landed once at most one younger group is pending
            if (np == 2) cp_async_wait<1>(); else cp_async_wait<0>();
            const WsMeta m = p_meta[head];
            const uint32_t st = p_stage[head], sb = smem_addr(srec) + st * (WS_BATCH * REC_BYTES);
            if (m.cnt) {
                const float tx0f = (float)((m.tile % v.gx) * TILE), ty0f = (float)((m.tile / v.gx) * TILE);
#pragma unroll
                for (int u = 0; u < WS_BATCH / 32; ++u) {
                    const uint32_t k = lane + 32u * u;
                    if (k < m.cnt) {
                        const float4 a = lds128(sb + k * REC_BYTES);
                        float4 c = lds128(sb + k * REC_BYTES + 16);
                        const uint32_t mk = block_mask8(a.x, a.y, a.z, c.x, c.y, c.z, tx0f, ty0f);
                        smask[st][k] = (uint8_t)mk;
                        masks[p_gbase[head] + k] = (uint8_t)mk;
                        c.x *= -0.5f * LOG2E, c.y *= -LOG2E, c.z *= -0.5f * LOG2E;
                        sts128(sb + k * REC_BYTES + 16, c);
                    }
                }
            }
            if (lane == 0) smeta[st] = m;
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[st]));
            head ^= 1, --np;
            if (m.tile < 0) break;
        }
        return;
    }

    // -------------------------------------------------- consumers --------------------------------------------------
    const uint32_t wbits = (1u << (4 * (warp >> 1) + (warp & 1))) | (1u << (4 * (warp >> 1) + 2 + (warp & 1)));
    const size_t hw = (size_t)v.H * v.W;
    int cur_tile = -1, px = 0, pya = 0, pyb = 0;
    bool ina = false, inb = false, tile_done = false;
    float pxf = 0.f, pyaf = 0.f;
    f2 T2 = bc(-1.0f), C0 = bc(0.f), C1 = bc(0.f), C2 = bc(0.f), D = bc(0.f);
    uint32_t lasta = 0, lastb = 0;
    for (;;) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        const WsMeta m = smeta[stage];
        if (m.tile < 0) break;
        if (m.tile != cur_tile) {  // first batch of a tile: this warp's 8x8 block, fresh pixel state
            cur_tile = m.tile, tile_done = false;
            const int bx0 = (m.tile % v.gx) * TILE + (warp & 1) * 8, by0 = (m.tile / v.gx) * TILE + (warp >> 1) * 8;
            px = bx0 + (lane & 7), pya = by0 + (lane >> 3), pyb = pya + 4;
            ina = px < v.W && pya < v.H, inb = px < v.W && pyb < v.H;
            pxf = (float)px, pyaf = (float)pya;
            T2 = pk(ina ? 1.0f : -1.0f, inb ? 1.0f : -1.0f), C0 = C1 = C2 = D = bc(0.f);
            lasta = lastb = 0;
        }
        const uint32_t cnt = tile_done ? 0u : m.cnt;
        const uint32_t sbase = smem_addr(srec) + stage * (WS_BATCH * REC_BYTES);
        if (cnt && !__all_sync(0xffffffffu, lo(T2) < 0.0f && hi(T2) < 0.0f)) {
            for (uint32_t r = 0; r < cnt; r += 32) {
                // lane l tests list entry r + 31 - l: bit b of the ballot is entry r + 31 - b, the highest bit comes first
                const uint32_t j = r + 31 - lane;
                const bool hit = j < cnt && (smask[stage][j] & wbits);
                uint32_t mask = __ballot_sync(0xffffffffu, hit);
                const uint32_t top = sbase + (r + 31) * REC_BYTES;  // record of bit 0 ... minus b records for bit b
                const uint32_t last_top = m.off + r + 32;           // 1-based list position of bit 0's entry ... - b
                while (mask) {
                    uint32_t b, src;  // opaque to the optimiser, which otherwise rebuilds the index from 31 - clz
                    asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(mask));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(src) : "r"(b), "r"(0u - (uint32_t)REC_BYTES), "r"(top));
                    mask ^= 1u << b;
                    const float2 xy = lds64(src);
                    const float4 c = lds128(src + 16);
                    const float4 col = lds128(src + 32);
                    const float dx = xy.x - pxf, dya = xy.y - pyaf;
                    const f2 dy2 = pk(dya, dya - 4.0f);
                    const f2 power2 = fma2(bc(dx), fma2(bc(c.x), bc(dx), mul2(bc(c.y), dy2)), mul2(mul2(bc(c.z), dy2), dy2));
                    const float pa = lo(power2), pb = hi(power2);
                    const f2 raw2 = mul2(bc(c.w), pk(ex2_approx(pa), ex2_approx(pb)));
                    const float ala = fminf(ALPHA_MAX, lo(raw2)), alb = fminf(ALPHA_MAX, hi(raw2));
                    const bool acta = (pa <= 0.0f) && (ala >= ALPHA_MIN), actb = (pb <= 0.0f) && (alb >= ALPHA_MIN);
                    const f2 al2 = pk(ala, alb);
                    const f2 Tn2 = mul2(T2, fma2(al2, bc(-1.0f), bc(1.0f)));  // T (1 - alpha)
                    const float Ta = lo(T2), Tb = hi(T2), Tna = lo(Tn2), Tnb = hi(Tn2);
                    const bool bla = acta && (Tna >= T_EPS), blb = actb && (Tnb >= T_EPS);
                    const f2 aT2 = mul2(al2, T2);
                    const f2 w2 = pk(bla ? lo(aT2) : 0.0f, blb ? hi(aT2) : 0.0f);
                    C0 = fma2(bc(col.x), w2, C0);
                    C1 = fma2(bc(col.y), w2, C1);
                    C2 = fma2(bc(col.z), w2, C2);
                    D = fma2(bc(col.w), w2, D);
                    const float Tsa = acta ? -fabsf(Ta) : Ta, Tsb = actb ? -fabsf(Tb) : Tb;
                    T2 = pk(bla ? Tna : Tsa, blb ? Tnb : Tsb);
                    const uint32_t pos = last_top - b;
                    lasta = bla ? pos : lasta;
                    lastb = blb ? pos : lastb;
                }
                if (__all_sync(0xffffffffu, lo(T2) < 0.0f && hi(T2) < 0.0f)) break;
            }
        }
        if (!m.last && !tile_done) {  // (uniform over the consumers) have all pixels of the tile terminated?
            tile_done = consumers_all(lo(T2) < 0.0f && hi(T2) < 0.0f);
            if (tile_done && tid == 0) *done_tile = m.tile;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
        if (++stage == (uint32_t)WS_STAGES) stage = 0, phase ^= 1u;
        if (m.last) {
            if (ina) {
                const float Tf = fabsf(lo(T2));
                const size_t pix = (size_t)pya * v.W + px;
                out_color[pix] = fmaf(Tf, v.bg[0], lo(C0));
                out_color[hw + pix] = fmaf(Tf, v.bg[1], lo(C1));
                out_color[2 * hw + pix] = fmaf(Tf, v.bg[2], lo(C2));
                out_depth[pix] = lo(D);
                final_T[pix] = Tf;
                n_contrib[pix] = lasta;
            }
            if (inb) {
                const float Tf = fabsf(hi(T2));
                const size_t pix = (size_t)pyb * v.W + px;
                out_color[pix] = fmaf(Tf, v.bg[0], hi(C0));
                out_color[hw + pix] = fmaf(Tf, v.bg[1], hi(C1));
                out_color[2 * hw + pix] = fmaf(Tf, v.bg[2], hi(C2));
                out_depth[pix] = hi(D);
                final_T[pix] = Tf;
                n_contrib[pix] = lastb;
            }
        }
    }
}

void launch_render_forward(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, uint32_t capacity, float* out_color,
                           float* out_depth, cudaStream_t s) {
#if GGRT_FWD_WS
    {
        static int sms = 0;  // persistent CTAs: as many as stay resident (per device the same on one box)
        if (sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        const int T = v.gx * v.gy, grid = T < sms * GGRT_FWD_WS_MINBLOCKS ? T : sms * GGRT_FWD_WS_MINBLOCKS;
        render_forward_ws_kernel<<<grid, WS_THREADS, 0, s>>>(v, g.rec0, g.rec1, g.rec2, im.starts, b.points, b.masks, capacity,
                                                             out_color, out_depth, im.final_T, im.n_contrib);
        return;
    }
#endif
#if GGRT_FWD_PAIR
    launch_chain(render_forward_pair_kernel, dim3(v.gx, v.gy, 1), dim3(F2_THREADS), 0, s, v, g.rec0, g.rec1, g.rec2, im.starts,
                 b.points, b.masks, capacity, out_color, out_depth, im.final_T, im.n_contrib);
    return;
#endif
    dim3 grid(v.gx, v.gy, 8 / FWD_WARPS);
    render_forward_kernel<GGRT_FWD_ASYNC != 0><<<grid, FWD_THREADS, 0, s>>>(v, g.rec0, g.rec1, g.rec2, im.starts, b.points,
                                                                             b.masks, capacity, out_color, out_depth,
                                                                             im.final_T, im.n_contrib);
}

}  // namespace ggrt
