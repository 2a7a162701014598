// Internal declarations shared by the kernels of libggrt_raster.so (sm_100a only).
// Algorithm spec: SURVEY.md Appendix A; call site cuda_splatting.py:101-125.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ggrt_raster.h"

namespace ggrt {

constexpr int TILE = GGRT_RASTER_TILE;
constexpr int TILE_PIXELS = TILE * TILE;
constexpr float NEAR_CULL = 0.2f;   // A.1 in_frustum threshold (stock 3DGS value)
constexpr float LOWPASS = 0.3f;     // A.1 screen-space dilation
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float T_EPS = 0.0001f;
constexpr float RADIUS_CAP = 1.0e6f;
constexpr int GRAD_STRIDE = 12;     // floats per Gaussian in the backward scratch
// Each tile owns SUBS pair counters: same-address atomics serialise in L2, so a tile's ~300 increments are spread
// over 16 addresses by the Gaussian index (idx mod 16), which cuts the binning kernels' critical path ~16x.
// The counters come in two banks of 16.  Bank 0 counts the pairs of Gaussians that touch at most RANKED_TILES tiles:
// the geometry kernel increments it with RETURNING atomics and keeps the returned ranks (GeomPtrs::ranks), so that
// `emit` places those pairs without any atomic (segment start + rank).  Bank 1 counts the pairs of larger Gaussians,
// which `emit` places with allocation atomics as before.  Sub-segments are contiguous inside the tile segment.
constexpr int SUBS = 32;
constexpr int SUBS_LOG2 = 5;
constexpr int SUB_LANES = 16;     // counters per bank
constexpr int RANKED_TILES = 4;   // Gaussians touching at most this many tiles get their slots ranked by geometry

// grad_scratch slot meaning (A.4): NDC-mean x,y | conic A, B(half convention), C | opacity | colour r,g,b
enum GradSlot { G_MX = 0, G_MY = 1, G_CA = 2, G_CB = 3, G_CC = 4, G_OP = 5, G_R = 6, G_G = 7, G_B = 8, G_AUX = 9 };

struct View {  // kernel-side copy of the per-call scalars (matrices stay in device memory)
    int W, H, gx, gy, P, deg, K;
    float tanfovx, tanfovy, fx, fy;
    float scale;        // scene scale applied to means (s) and covariances (s*s)
    int cov_stride;     // 6 ([P,6]) or 9 ([P,3,3])
    int sh_ks, sh_cs;   // SH element (k, c) of a Gaussian's row lives at k*sh_ks + c*sh_cs
    const float* view;
    const float* proj;
    const float* campos;
    const float* bg;
    const float* dparams;  // device {tanfovx, tanfovy, scene_scale} overriding the host copies above (or NULL)
    int aux_mode;          // 1: aux channel = max(0, C0 z + 0.5) of the unscaled view depth (GGRt's depth pass)
};

struct GeomPtrs {
    float4* rec0;      // {pix_x, pix_y, cull threshold, view depth}
    float4* rec1;      // {conic A, B, C, opacity}
    float4* rec2;      // {r, g, b, aux}
    ushort4* rect;     // tile rect
    uint32_t* tiles;   // tiles touched
    uint8_t* flags;    // clamp bits
    uint4* ranks;      // per Gaussian touching <= RANKED_TILES tiles: its rank in the (tile, sub-counter) segment of each
    float* jac;        // 9 planes of P floats: d(SH colour c)/d(scaled mean j) at plane 3c + j, written by the colour kernel
                       // (which has the SH rows in shared memory anyway) so that the backward never re-reads the SH table
    size_t jac_plane;  // floats per plane
};

#ifndef GGRT_SCAN_BLOCK
#define GGRT_SCAN_BLOCK 256
#endif
constexpr int SCAN_BLOCK = GGRT_SCAN_BLOCK;  // tiles per CTA of the tile scan

struct ImagePtrs {
    uint32_t* counts;
    unsigned long long* partials;
    uint32_t* starts;
    uint32_t* cursor;
    uint32_t* header;
    float* final_T;
    uint32_t* n_contrib;
};

struct BinPtrs {
    unsigned long long* keys;
    uint32_t* points;
    uint8_t* masks;  // per list entry: which of the tile's 8 warp pixel blocks the Gaussian reaches (render kernels)
};

// destinations of the compact colour gradients: [P,3] rows of dL/drgb, plus (with_campos) a row P holding the
// view's camera centre, so that a peer can rebuild dL/dsh from the buffer alone
struct ColorSinks {
    float* ptr[GGRT_RASTER_MAX_MERGE_VIEWS];
    int n;
    int multimem;     // ptr[0] is an NVLS multicast address: use multimem.st
    int with_campos;
    // step signalling (GgrtRasterGradSinks.epoch): see include/ggrt_raster.h
    uint32_t* epoch;
    uint32_t* done;
    long long parity_stride;
    int n_arrive;
    uint32_t* arrive[GGRT_RASTER_MAX_MERGE_VIEWS];
};

// ---- cross-GPU signalling primitives (system scope) -------------------------------------------------------------
__device__ __forceinline__ void signal_add(uint32_t* counter, bool multimem) {
    if (multimem)
        asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    else
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
}
__device__ __forceinline__ void wait_reached(const uint32_t* counter, uint32_t target) {
    uint32_t v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);  // wrap-around safe
}
// Called by every thread of every CTA at the end of a kernel whose global stores must be published to other GPUs:
// returns true in thread 0 of the LAST CTA to get here, after all CTAs' stores have been fenced at system scope.
__device__ __forceinline__ bool last_cta_done(uint32_t* done_counter) {
    __syncthreads();
    if (threadIdx.x != 0) return false;
    __threadfence_system();
    const uint32_t total = gridDim.x * gridDim.y * gridDim.z;
    if (atomicAdd(done_counter, 1u) != total - 1u) return false;
    *done_counter = 0u;  // ready for the next launch
    __threadfence_system();
    return true;
}

// ---- programmatic dependent launch (PDL) along the chain scan -> emit -> sort -> render fwd -> render bwd -> per-Gaussian
// bwd.  Each of these kernels calls pdl_enter() before its first global access: it lets the NEXT kernel of the stream be
// scheduled as soon as every CTA of this one has started (the dependents then sit in freed SM slots during this kernel's
// tail), and waits until the PREVIOUS kernel has completed and its memory is visible.  launch_chain() launches with
// programmatic stream serialisation, so the launch latency and the CTA ramp of each kernel hide under its predecessor;
// without the attribute both instructions are no-ops.  The attribute is only used while the stream is being CAPTURED
// into a CUDA graph (graph.CapturedStep / CapturedViews, where it becomes a programmatic graph edge: 0.3030 -> 0.3001 ms
// per step at C2): on a live stream that also carries the side-stream event waits, launches with the attribute cost the
// HOST 0.33 ms more per frame (measured: decoder-level call 0.87 -> 1.20 ms), and an eagerly launched frame is host
// bound already.  GGRT_RASTER_PDL=0 switches it off everywhere.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled(cudaStream_t s);
template <typename... KArgs, typename... Args>
inline void launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = pdl_enabled(s) ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

void compute_layout(int P, int H, int W, long long N, GgrtRasterLayout* L);
GeomPtrs geom_ptrs(void* base, int P);
ImagePtrs image_ptrs(void* base, int H, int W);
BinPtrs bin_ptrs(void* base, long long N);

// kernel launchers (each in its own translation unit)
void launch_geometry(const View& v, const float* means, const float* cov3d, const float* opac, int* radii,
                     GeomPtrs g, ImagePtrs im, cudaStream_t s);
void launch_scan_tiles(const View& v, ImagePtrs im, uint32_t* counts_host, cudaStream_t s);
void launch_color(const View& v, const float* means, const float* shs, const float* colors, const float* aux,
                  const int* radii, GeomPtrs g, cudaStream_t s);
void launch_emit(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, uint32_t capacity, cudaStream_t s);
void launch_sort_tiles(const View& v, ImagePtrs im, BinPtrs b, uint32_t max_tile_pairs, uint32_t capacity,
                       cudaStream_t s);
void launch_render_forward(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, uint32_t capacity, float* out_color,
                           float* out_depth, cudaStream_t s);
void launch_render_backward(const View& v, GeomPtrs g, ImagePtrs im, BinPtrs b, const float* dL_dout,
                            const float* dL_dout_aux, float* scratch, cudaStream_t s);
void launch_preprocess_backward(const View& v, const float* means, const float* cov3d, const float* shs,
                                const int* radii, GeomPtrs g, const float* scratch, float* dmeans2D, float* dopacity,
                                float* dmeans3D, float* dcov3D, float* dsh, float* dcolors, float* daux, float* dcam,
                                const ColorSinks& sinks, cudaStream_t s, cudaStream_t sh_stream = nullptr);
struct MergeSignal {  // signalled exchange: wait for world * *epoch arrivals, read half (*epoch - 1) & 1
    const uint32_t* epoch = nullptr;
    const uint32_t* arrive = nullptr;
    int world = 0;
    long long parity_stride = 0;
};
void launch_sh_gradient_merge(int P, int deg, float scale, bool cmajor, const float* means, int num_views,
                              const float* const* drgb, const float* const* campos, float* dsh, const MergeSignal& sig,
                              cudaStream_t s);
void launch_nvls_allreduce(float* multicast, long long count, int rank, int world, const uint32_t* epoch,
                           const uint32_t* arrive_in, uint32_t* arrive_out_mc, const uint32_t* arrive_out,
                           uint32_t* done_counter, cudaStream_t s);
void launch_nvls_barrier(unsigned int* mc_counter, const unsigned int* local_counter, unsigned int target,
                         cudaStream_t s);
void launch_adapter_forward(const GgrtAdapterParams& p, const float* extr, const float* intr, const float* shrot,
                            const float* coords, const float* depths, const float* raw, float* means, float* cov,
                            float* harm, float* scales_out, float* rot_out, cudaStream_t s);
void launch_adapter_backward(const GgrtAdapterParams& p, const float* extr, const float* intr, const float* shrot,
                             const float* coords, const float* depths, const float* raw, const float* dmeans,
                             const float* dcov, const float* dharm, float* dcoords, float* ddepths, float* draw,
                             cudaStream_t s);
void launch_mark_visible(int P, const float* means, const float* view, uint8_t* present, cudaStream_t s);

// ---------------------------------------------------------------------------------------
// Exactly-rounded float helpers.  The geometry path (cull, cov2D, radius, tile rect,
// depth key) must be bit-identical to the CPU oracle, which is compiled without FMA
// contraction; these intrinsics are never fused by the compiler.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
// (a*b + c*d) + e*f with separate roundings, left to right
__device__ __forceinline__ float dot3(float a, float b, float c, float d, float e, float f) {
    return fadd(fadd(fmul(a, b), fmul(c, d)), fmul(e, f));
}

struct Geo {
    float tx, ty, tz, cx, cy, txtz, tytz, hx, hy, hw, pw;
    float Tm[2][3];
    float a, b, c, det;
};

// A.1 geometric core; returns false if culled.  Same operation order as oracle/raster_oracle.c:geometry().
__device__ __forceinline__ bool geometry(const View& v, const float* __restrict__ V, const float* __restrict__ M,
                                         float px, float py, float pz, const float cv[6], Geo& g) {
    g.tx = fadd(dot3(V[0], px, V[4], py, V[8], pz), V[12]);
    g.ty = fadd(dot3(V[1], px, V[5], py, V[9], pz), V[13]);
    g.tz = fadd(dot3(V[2], px, V[6], py, V[10], pz), V[14]);
    if (!(g.tz > NEAR_CULL)) return false;
    g.hx = fadd(dot3(M[0], px, M[4], py, M[8], pz), M[12]);
    g.hy = fadd(dot3(M[1], px, M[5], py, M[9], pz), M[13]);
    g.hw = fadd(dot3(M[3], px, M[7], py, M[11], pz), M[15]);
    g.pw = fdiv(1.0f, fadd(g.hw, 0.0000001f));
    const float limx = fmul(1.3f, v.tanfovx), limy = fmul(1.3f, v.tanfovy);
    g.txtz = fdiv(g.tx, g.tz);
    g.tytz = fdiv(g.ty, g.tz);
    g.cx = fmul(fminf(limx, fmaxf(-limx, g.txtz)), g.tz);
    g.cy = fmul(fminf(limy, fmaxf(-limy, g.tytz)), g.tz);
    const float tz2 = fmul(g.tz, g.tz);
    const float J00 = fdiv(v.fx, g.tz), J02 = fdiv(-fmul(v.fx, g.cx), tz2);
    const float J11 = fdiv(v.fy, g.tz), J12 = fdiv(-fmul(v.fy, g.cy), tz2);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float r0 = V[4 * k + 0], r1 = V[4 * k + 1], r2 = V[4 * k + 2];
        g.Tm[0][k] = fadd(fmul(J00, r0), fmul(J02, r2));
        g.Tm[1][k] = fadd(fmul(J11, r1), fmul(J12, r2));
    }
    const float s00 = cv[0], s01 = cv[1], s02 = cv[2], s11 = cv[3], s12 = cv[4], s22 = cv[5];
    const float* t0 = g.Tm[0];
    const float* t1 = g.Tm[1];
    const float v00 = dot3(s00, t0[0], s01, t0[1], s02, t0[2]);
    const float v01 = dot3(s01, t0[0], s11, t0[1], s12, t0[2]);
    const float v02 = dot3(s02, t0[0], s12, t0[1], s22, t0[2]);
    const float v10 = dot3(s00, t1[0], s01, t1[1], s02, t1[2]);
    const float v11 = dot3(s01, t1[0], s11, t1[1], s12, t1[2]);
    const float v12 = dot3(s02, t1[0], s12, t1[1], s22, t1[2]);
    g.a = fadd(dot3(t0[0], v00, t0[1], v01, t0[2], v02), LOWPASS);
    g.b = dot3(t1[0], v00, t1[1], v01, t1[2], v02);
    g.c = fadd(dot3(t1[0], v10, t1[1], v11, t1[2], v12), LOWPASS);
    g.det = fsub(fmul(g.a, g.c), fmul(g.b, g.b));
    if (!(g.det > 0.0f) && !(g.det < 0.0f)) return false;  // det == 0 (A.1) or NaN
    return true;
}

// the 6 upper-triangular covariance entries (xx,xy,xz,yy,yz,zz) of Gaussian i, as stored (prefetchable: no use
// of the loaded values); scale_cov6 applies the scene scale afterwards
__device__ __forceinline__ void load_cov6_raw(const View& v, const float* __restrict__ cov3d, int i, float cv[6]) {
    const int stride = v.cov_stride;
    const float* c = cov3d + (size_t)i * stride;
    if (stride == 9) {
        cv[0] = c[0], cv[1] = c[1], cv[2] = c[2], cv[3] = c[4], cv[4] = c[5], cv[5] = c[8];
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) cv[k] = c[k];
    }
}
// s*s scaling, exactly rounded, matching the reference's separate `covariances * scale**2` (cuda_splatting.py:70)
__device__ __forceinline__ void scale_cov6(const View& v, float cv[6]) {
    const float s2 = fmul(v.scale, v.scale);
#pragma unroll
    for (int k = 0; k < 6; ++k) cv[k] = fmul(cv[k], s2);
}
__device__ __forceinline__ void load_cov6(const View& v, const float* __restrict__ cov3d, int i, float cv[6]) {
    const int stride = v.cov_stride;
    const float* c = cov3d + (size_t)i * stride;
    if (stride == 9) {
        cv[0] = c[0], cv[1] = c[1], cv[2] = c[2], cv[3] = c[4], cv[4] = c[5], cv[5] = c[8];
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) cv[k] = c[k];
    }
    const float s2 = fmul(v.scale, v.scale);
#pragma unroll
    for (int k = 0; k < 6; ++k) cv[k] = fmul(cv[k], s2);
}

__device__ __forceinline__ int f2i_sat(float x) {
    x = fmaxf(x, -1.0f);
    x = fminf(x, 65536.0f);
    return (int)x;
}

// SH constants (A.1)
#define GGRT_SH_C0 0.28209479177387814f
#define GGRT_SH_C1 0.4886025119029199f
#define GGRT_SH_C2_0 1.0925484305920792f
#define GGRT_SH_C2_1 -1.0925484305920792f
#define GGRT_SH_C2_2 0.31539156525252005f
#define GGRT_SH_C2_3 -1.0925484305920792f
#define GGRT_SH_C2_4 0.5462742152960396f
#define GGRT_SH_C3_0 -0.5900435899266435f
#define GGRT_SH_C3_1 2.890611442640554f
#define GGRT_SH_C3_2 -0.4570457994644658f
#define GGRT_SH_C3_3 0.3731763325901154f
#define GGRT_SH_C3_4 -0.4570457994644658f
#define GGRT_SH_C3_5 1.445305721320277f
#define GGRT_SH_C3_6 -0.5900435899266435f
#define GGRT_SH_C4_0 2.5033429417967046f
#define GGRT_SH_C4_1 -1.7701307697799304f
#define GGRT_SH_C4_2 0.9461746957575601f
#define GGRT_SH_C4_3 -0.6690465435572892f
#define GGRT_SH_C4_4 0.10578554691520431f
#define GGRT_SH_C4_5 -0.6690465435572892f
#define GGRT_SH_C4_6 0.47308734787878004f
#define GGRT_SH_C4_7 -1.7701307697799304f
#define GGRT_SH_C4_8 0.6258357354491761f

// The real SH basis up to degree 4 as a list of terms T(k, B_k, dB_k/dx, dB_k/dy, dB_k/dz) at the direction (x, y, z)
// (derivatives of the polynomials; the normalisation of the direction is differentiated separately).  The users define
// T and need xx, yy, zz, xy, yz, xz in scope; B_k is the same expression as sh_basis() below.
#define GGRT_SH_TERMS_0(T) T(0, GGRT_SH_C0, GGRT_Z, GGRT_Z, GGRT_Z)
#define GGRT_SH_TERMS_1(T)                                   \
    T(1, -GGRT_SH_C1 * y, GGRT_Z, -GGRT_SH_C1, GGRT_Z)             \
    T(2, GGRT_SH_C1 * z, GGRT_Z, GGRT_Z, GGRT_SH_C1)               \
    T(3, -GGRT_SH_C1 * x, -GGRT_SH_C1, GGRT_Z, GGRT_Z)
#define GGRT_SH_TERMS_2(T)                                                                                         \
    T(4, GGRT_SH_C2_0 * xy, GGRT_SH_C2_0 * y, GGRT_SH_C2_0 * x, GGRT_Z)                                              \
    T(5, GGRT_SH_C2_1 * yz, GGRT_Z, GGRT_SH_C2_1 * z, GGRT_SH_C2_1 * y)                                              \
    T(6, GGRT_SH_C2_2 * (2.0f * zz - xx - yy), GGRT_SH_C2_2 * -2.0f * x, GGRT_SH_C2_2 * -2.0f * y,                \
      GGRT_SH_C2_2 * 4.0f * z)                                                                                    \
    T(7, GGRT_SH_C2_3 * xz, GGRT_SH_C2_3 * z, GGRT_Z, GGRT_SH_C2_3 * x)                                              \
    T(8, GGRT_SH_C2_4 * (xx - yy), GGRT_SH_C2_4 * 2.0f * x, GGRT_SH_C2_4 * -2.0f * y, GGRT_Z)
#define GGRT_SH_TERMS_3(T)                                                                                         \
    T(9, GGRT_SH_C3_0 * y * (3.0f * xx - yy), GGRT_SH_C3_0 * 6.0f * xy, GGRT_SH_C3_0 * (3.0f * xx - 3.0f * yy), GGRT_Z) \
    T(10, GGRT_SH_C3_1 * xy * z, GGRT_SH_C3_1 * yz, GGRT_SH_C3_1 * xz, GGRT_SH_C3_1 * xy)                          \
    T(11, GGRT_SH_C3_2 * y * (4.0f * zz - xx - yy), GGRT_SH_C3_2 * -2.0f * xy,                                    \
      GGRT_SH_C3_2 * (4.0f * zz - xx - 3.0f * yy), GGRT_SH_C3_2 * 8.0f * yz)                                      \
    T(12, GGRT_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy), GGRT_SH_C3_3 * -6.0f * xz,                      \
      GGRT_SH_C3_3 * -6.0f * yz, GGRT_SH_C3_3 * (6.0f * zz - 3.0f * xx - 3.0f * yy))                              \
    T(13, GGRT_SH_C3_4 * x * (4.0f * zz - xx - yy), GGRT_SH_C3_4 * (4.0f * zz - 3.0f * xx - yy),                  \
      GGRT_SH_C3_4 * -2.0f * xy, GGRT_SH_C3_4 * 8.0f * xz)                                                        \
    T(14, GGRT_SH_C3_5 * z * (xx - yy), GGRT_SH_C3_5 * 2.0f * xz, GGRT_SH_C3_5 * -2.0f * yz,                      \
      GGRT_SH_C3_5 * (xx - yy))                                                                                   \
    T(15, GGRT_SH_C3_6 * x * (xx - 3.0f * yy), GGRT_SH_C3_6 * (3.0f * xx - 3.0f * yy), GGRT_SH_C3_6 * -6.0f * xy, GGRT_Z)
#define GGRT_SH_TERMS_4(T)                                                                                         \
    T(16, GGRT_SH_C4_0 * xy * (xx - yy), GGRT_SH_C4_0 * y * (3.0f * xx - yy), GGRT_SH_C4_0 * x * (xx - 3.0f * yy), GGRT_Z) \
    T(17, GGRT_SH_C4_1 * yz * (3.0f * xx - yy), GGRT_SH_C4_1 * 6.0f * xy * z,                                     \
      GGRT_SH_C4_1 * z * (3.0f * xx - 3.0f * yy), GGRT_SH_C4_1 * y * (3.0f * xx - yy))                            \
    T(18, GGRT_SH_C4_2 * xy * (7.0f * zz - 1.0f), GGRT_SH_C4_2 * y * (7.0f * zz - 1.0f),                          \
      GGRT_SH_C4_2 * x * (7.0f * zz - 1.0f), GGRT_SH_C4_2 * 14.0f * xy * z)                                       \
    T(19, GGRT_SH_C4_3 * yz * (7.0f * zz - 3.0f), GGRT_Z, GGRT_SH_C4_3 * z * (7.0f * zz - 3.0f),                     \
      GGRT_SH_C4_3 * y * (21.0f * zz - 3.0f))                                                                     \
    T(20, GGRT_SH_C4_4 * (zz * (35.0f * zz - 30.0f) + 3.0f), GGRT_Z, GGRT_Z, GGRT_SH_C4_4 * (140.0f * zz * z - 60.0f * z)) \
    T(21, GGRT_SH_C4_5 * xz * (7.0f * zz - 3.0f), GGRT_SH_C4_5 * z * (7.0f * zz - 3.0f), GGRT_Z,                     \
      GGRT_SH_C4_5 * x * (21.0f * zz - 3.0f))                                                                     \
    T(22, GGRT_SH_C4_6 * (xx - yy) * (7.0f * zz - 1.0f), GGRT_SH_C4_6 * 2.0f * x * (7.0f * zz - 1.0f),            \
      GGRT_SH_C4_6 * -2.0f * y * (7.0f * zz - 1.0f), GGRT_SH_C4_6 * (xx - yy) * 14.0f * z)                        \
    T(23, GGRT_SH_C4_7 * xz * (xx - 3.0f * yy), GGRT_SH_C4_7 * z * (3.0f * xx - 3.0f * yy),                       \
      GGRT_SH_C4_7 * -6.0f * xy * z, GGRT_SH_C4_7 * x * (xx - 3.0f * yy))                                         \
    T(24, GGRT_SH_C4_8 * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy)),                                         \
      GGRT_SH_C4_8 * (4.0f * xx * x - 12.0f * x * yy), GGRT_SH_C4_8 * (-12.0f * xx * y + 4.0f * yy * y), GGRT_Z)
// acc += coefficient * s; GGRT_Z marks the identically-zero derivatives of the term list, whose FMA is dropped at
// compile time by overload resolution (fmaf(0, s, acc) itself could not be: s may be Inf / NaN)
struct ShZero {};
#define GGRT_Z (ShZero{})
__device__ __forceinline__ void sh_fma(ShZero, float, float&) {}
__device__ __forceinline__ void sh_fma(float coef, float s, float& acc) { acc = fmaf(coef, s, acc); }

// Real SH basis of degree `deg` at unit direction (x,y,z) -> b[0..K)
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
    b[0] = GGRT_SH_C0;
    if (deg < 1) return;
    b[1] = -GGRT_SH_C1 * y;
    b[2] = GGRT_SH_C1 * z;
    b[3] = -GGRT_SH_C1 * x;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = GGRT_SH_C2_0 * xy;
    b[5] = GGRT_SH_C2_1 * yz;
    b[6] = GGRT_SH_C2_2 * (2.0f * zz - xx - yy);
    b[7] = GGRT_SH_C2_3 * xz;
    b[8] = GGRT_SH_C2_4 * (xx - yy);
    if (deg < 3) return;
    b[9] = GGRT_SH_C3_0 * y * (3.0f * xx - yy);
    b[10] = GGRT_SH_C3_1 * xy * z;
    b[11] = GGRT_SH_C3_2 * y * (4.0f * zz - xx - yy);
    b[12] = GGRT_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[13] = GGRT_SH_C3_4 * x * (4.0f * zz - xx - yy);
    b[14] = GGRT_SH_C3_5 * z * (xx - yy);
    b[15] = GGRT_SH_C3_6 * x * (xx - 3.0f * yy);
    if (deg < 4) return;
    b[16] = GGRT_SH_C4_0 * xy * (xx - yy);
    b[17] = GGRT_SH_C4_1 * yz * (3.0f * xx - yy);
    b[18] = GGRT_SH_C4_2 * xy * (7.0f * zz - 1.0f);
    b[19] = GGRT_SH_C4_3 * yz * (7.0f * zz - 3.0f);
    b[20] = GGRT_SH_C4_4 * (zz * (35.0f * zz - 30.0f) + 3.0f);
    b[21] = GGRT_SH_C4_5 * xz * (7.0f * zz - 3.0f);
    b[22] = GGRT_SH_C4_6 * (xx - yy) * (7.0f * zz - 1.0f);
    b[23] = GGRT_SH_C4_7 * xz * (xx - 3.0f * yy);
    b[24] = GGRT_SH_C4_8 * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
}

// Kernels whose maths uses the camera scalars call this first: with a device-resident camera (View::dparams) the
// host never saw tanfov / scene scale, so they are read here; fx, fy are derived exactly as the host derives them.
__device__ __forceinline__ void resolve_device_params(View& v) {
    if (v.dparams != nullptr) {
        v.tanfovx = v.dparams[0], v.tanfovy = v.dparams[1], v.scale = v.dparams[2];
        v.fx = fdiv((float)v.W, fmul(2.0f, v.tanfovx));
        v.fy = fdiv((float)v.H, fmul(2.0f, v.tanfovy));
    }
}
// GGRt's depth channel: the degree-0 SH "colour" of the unscaled camera-space depth (cuda_splatting.py:256-268)
__device__ __forceinline__ float ggrt_depth_channel(float tz, float scale) {
    return fmaxf(0.0f, fadd(fmul(GGRT_SH_C0, fdiv(tz, scale)), 0.5f));
}

void launch_camera_setup(int n, const float* extr, const float* intr, const float* near, const float* far,
                         int scale_invariant, float* out, cudaStream_t s);
void set_error(const char* fmt, ...);
int check_launch(const char* what, int debug, cudaStream_t s);

}  // namespace ggrt
