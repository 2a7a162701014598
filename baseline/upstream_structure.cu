// "Upstream-structure" GPU baseline (baseline (ii)): the kernel STRUCTURE of the public 3DGS tile
// rasterizer that GGRt's external `diff_gaussian_rasterization` package implements -- per-Gaussian prefix sum of
// tiles_touched + a host read of N, duplicateWithKeys, ONE global 64-bit radix sort of all N (key, value) pairs
// (CUB), identifyTileRanges, a 16x16-thread render kernel per tile with cooperative fetches and no per-warp cull,
// and a render backward in which every pixel thread issues ~10 global float atomics per contributing Gaussian --
// restated from the published algorithm and compiled for sm_100a, so that the B200-native
// design can be timed against the stock structure ON THE SAME GPU.
//
// THIS IS NOT THE REFERENCE BINARY: the reference's rasterizer source is an un-vendored third-party package that
// is not available offline.  It is a measurement aid only: nothing in the product imports
// it.  To keep the comparison conservative, the per-Gaussian streaming kernels (geometry, SH colour, preprocess
// backward) are SHARED with the product library (upstream's are simpler and slower); only the binning and the
// two render kernels -- where the designs differ -- are restated here.
#include <cub/cub.cuh>

#include <cstdio>

#include "../ggrt_official_b200/csrc/common.cuh"

using namespace ggrt;

namespace {

constexpr int UB = 256;  // threads per tile (16 x 16)

struct Work {
    int P = 0, H = 0, W = 0, T = 0;
    size_t pair_cap = 0;
    void *geom = nullptr, *image = nullptr;
    uint32_t* offsets = nullptr;
    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    uint32_t *vals = nullptr, *vals_sorted = nullptr;
    uint2* ranges = nullptr;
    float* final_T = nullptr;
    uint32_t* n_contrib = nullptr;
    float* scratch = nullptr;
    void *scan_tmp = nullptr, *sort_tmp = nullptr;
    size_t scan_bytes = 0, sort_bytes = 0;
    uint32_t N = 0;
};
Work g_w;
char g_msg[256] = "";

#define UP_CHECK(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            snprintf(g_msg, sizeof(g_msg), "%s: %s", #call, cudaGetErrorString(e_));     \
            return -2;                                                                   \
        }                                                                                \
    } while (0)

int ensure_sizes(int P, int H, int W) {
    const int T = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    if (P == g_w.P && H == g_w.H && W == g_w.W) return 0;
    GgrtRasterLayout L;
    compute_layout(P, H, W, 0, &L);
    cudaFree(g_w.geom), cudaFree(g_w.image), cudaFree(g_w.offsets), cudaFree(g_w.ranges), cudaFree(g_w.final_T);
    cudaFree(g_w.n_contrib), cudaFree(g_w.scratch), cudaFree(g_w.scan_tmp);
    UP_CHECK(cudaMalloc(&g_w.geom, L.geom_bytes));
    UP_CHECK(cudaMalloc(&g_w.image, L.img_bytes));
    UP_CHECK(cudaMalloc(&g_w.offsets, sizeof(uint32_t) * (size_t)(P + 1)));
    UP_CHECK(cudaMalloc(&g_w.ranges, sizeof(uint2) * (size_t)T));
    UP_CHECK(cudaMalloc(&g_w.final_T, sizeof(float) * (size_t)H * W));
    UP_CHECK(cudaMalloc(&g_w.n_contrib, sizeof(uint32_t) * (size_t)H * W));
    UP_CHECK(cudaMalloc(&g_w.scratch, sizeof(float) * (size_t)P * GRAD_STRIDE));
    g_w.scan_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, g_w.scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
    UP_CHECK(cudaMalloc(&g_w.scan_tmp, g_w.scan_bytes + 256));
    g_w.P = P, g_w.H = H, g_w.W = W, g_w.T = T;
    return 0;
}

int ensure_pairs(size_t N) {
    if (N <= g_w.pair_cap) return 0;
    const size_t cap = N + N / 4 + 1024;
    cudaFree(g_w.keys), cudaFree(g_w.keys_sorted), cudaFree(g_w.vals), cudaFree(g_w.vals_sorted), cudaFree(g_w.sort_tmp);
    UP_CHECK(cudaMalloc(&g_w.keys, sizeof(unsigned long long) * cap));
    UP_CHECK(cudaMalloc(&g_w.keys_sorted, sizeof(unsigned long long) * cap));
    UP_CHECK(cudaMalloc(&g_w.vals, sizeof(uint32_t) * cap));
    UP_CHECK(cudaMalloc(&g_w.vals_sorted, sizeof(uint32_t) * cap));
    g_w.sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, g_w.sort_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int)cap);
    UP_CHECK(cudaMalloc(&g_w.sort_tmp, g_w.sort_bytes + 256));
    g_w.pair_cap = cap;
    return 0;
}

// one (tile << 32 | depth bits, gaussian) pair per touched tile, at the Gaussian's slots of the prefix sum (A.2)
__global__ void duplicate_with_keys(int P, int gx, const uint32_t* __restrict__ offsets, const ushort4* __restrict__ rect,
                                    const uint32_t* __restrict__ tiles, const float4* __restrict__ rec0,
                                    unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || tiles[i] == 0) return;
    uint32_t off = i == 0 ? 0u : offsets[i - 1];
    const ushort4 r = rect[i];
    const unsigned long long depth_bits = __float_as_uint(rec0[i].w);
    for (int y = r.y; y < r.w; ++y)
        for (int x = r.x; x < r.z; ++x) {
            keys[off] = ((unsigned long long)(uint32_t)(y * gx + x) << 32) | depth_bits;
            vals[off] = (uint32_t)i;
            ++off;
        }
}

__global__ void identify_tile_ranges(uint32_t N, const unsigned long long* __restrict__ keys, uint2* __restrict__ ranges) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t tile = (uint32_t)(keys[i] >> 32);
    if (i == 0)
        ranges[tile].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (tile != prev) {
            ranges[prev].y = i;
            ranges[tile].x = i;
        }
    }
    if (i == N - 1) ranges[tile].y = N;
}

// A.3: one thread per pixel, the 256 threads of a tile fetch 256 list entries at a time
__global__ void __launch_bounds__(UB)
render_forward_classic(int W, int H, int gx, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const float4* __restrict__ rec0, const float4* __restrict__ rec1, const float4* __restrict__ rec2,
                       const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_depth,
                       float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
    __shared__ uint32_t c_id[UB];
    __shared__ float2 c_xy[UB];
    __shared__ float4 c_co[UB];
    const int tx = threadIdx.x, ty = threadIdx.y, rank = ty * TILE + tx;
    const int px = blockIdx.x * TILE + tx, py = blockIdx.y * TILE + ty;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = (int)((range.y - range.x + UB - 1) / UB);
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t contributor = 0, last = 0;
    for (int i = 0; i < rounds; ++i, todo -= UB) {
        if (__syncthreads_count(done) == UB) break;
        const uint32_t progress = (uint32_t)i * UB + rank;
        if (range.x + progress < range.y) {
            const uint32_t id = point_list[range.x + progress];
            const float4 a = rec0[id];
            c_id[rank] = id;
            c_xy[rank] = make_float2(a.x, a.y);
            c_co[rank] = rec1[id];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(UB, todo); ++j) {
            ++contributor;
            const float dx = c_xy[j].x - pxf, dy = c_xy[j].y - pyf;
            const float4 co = c_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(ALPHA_MAX, co.w * __expf(power));
            if (alpha < ALPHA_MIN) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < T_EPS) {
                done = true;
                continue;
            }
            const float4 col = rec2[c_id[j]];
            const float w = alpha * T;
            C0 += col.x * w, C1 += col.y * w, C2 += col.z * w, D += col.w * w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px, hw = (size_t)H * W;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_color[pix] = C0 + T * bg[0];
        out_color[hw + pix] = C1 + T * bg[1];
        out_color[2 * hw + pix] = C2 + T * bg[2];
        out_depth[pix] = D;
    }
}

// A.4: back to front, every pixel thread commits its contribution with global float atomics
__global__ void __launch_bounds__(UB)
render_backward_classic(int W, int H, int gx, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                        const float4* __restrict__ rec0, const float4* __restrict__ rec1, const float4* __restrict__ rec2,
                        const float* __restrict__ bg, const float* __restrict__ final_T,
                        const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                        float* __restrict__ scratch) {
    __shared__ uint32_t c_id[UB];
    __shared__ float2 c_xy[UB];
    __shared__ float4 c_co[UB];
    __shared__ float4 c_col[UB];
    const int tx = threadIdx.x, ty = threadIdx.y, rank = ty * TILE + tx;
    const int px = blockIdx.x * TILE + tx, py = blockIdx.y * TILE + ty;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = (int)((range.y - range.x + UB - 1) / UB);
    int todo = (int)(range.y - range.x);
    const size_t pix = (size_t)py * W + px, hw = (size_t)H * W;
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    uint32_t contributor = (uint32_t)todo;
    const uint32_t last_contributor = inside ? n_contrib[pix] : 0u;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (inside) g0 = dL_dpix[pix], g1 = dL_dpix[hw + pix], g2 = dL_dpix[2 * hw + pix];
    const float bg_dot = bg[0] * g0 + bg[1] * g1 + bg[2] * g2;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;
    for (int i = 0; i < rounds; ++i, todo -= UB) {
        __syncthreads();
        const uint32_t progress = (uint32_t)i * UB + rank;
        if (range.x + progress < range.y) {
            const uint32_t id = point_list[range.y - progress - 1];
            const float4 a = rec0[id];
            c_id[rank] = id;
            c_xy[rank] = make_float2(a.x, a.y);
            c_co[rank] = rec1[id];
            c_col[rank] = rec2[id];
        }
        __syncthreads();
        for (int j = 0; inside && j < min(UB, todo); ++j) {
            --contributor;
            if (contributor >= last_contributor) continue;
            const float dx = c_xy[j].x - pxf, dy = c_xy[j].y - pyf;
            const float4 co = c_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            const float G = __expf(power);
            const float alpha = fminf(ALPHA_MAX, co.w * G);
            if (alpha < ALPHA_MIN) continue;
            T = T / (1.0f - alpha);
            const float w = alpha * T;
            float* dst = scratch + (size_t)c_id[j] * GRAD_STRIDE;
            const float4 c = c_col[j];
            float dL_dalpha = 0.f;
            acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0, lc0 = c.x, dL_dalpha += (c.x - acc0) * g0;
            acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1, lc1 = c.y, dL_dalpha += (c.y - acc1) * g1;
            acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2, lc2 = c.z, dL_dalpha += (c.z - acc2) * g2;
            atomicAdd(dst + G_R, w * g0);
            atomicAdd(dst + G_G, w * g1);
            atomicAdd(dst + G_B, w * g2);
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            atomicAdd(dst + G_MX, dL_dG * (-gdx * co.x - gdy * co.y) * ddelx_dx);
            atomicAdd(dst + G_MY, dL_dG * (-gdy * co.z - gdx * co.y) * ddely_dy);
            atomicAdd(dst + G_CA, -0.5f * gdx * dx * dL_dG);
            atomicAdd(dst + G_CB, -0.5f * gdx * dy * dL_dG);
            atomicAdd(dst + G_CC, -0.5f * gdy * dy * dL_dG);
            atomicAdd(dst + G_OP, G * dL_dalpha);
        }
    }
}

int make_view(const GgrtRasterSettings* s, int P, View* v) {
    *v = View{};  // dparams = NULL, aux_mode = 0: host-side scalars, plain depth channel
    v->W = s->image_width, v->H = s->image_height;
    v->gx = (v->W + TILE - 1) / TILE, v->gy = (v->H + TILE - 1) / TILE;
    v->P = P, v->deg = s->sh_degree, v->K = (s->sh_degree + 1) * (s->sh_degree + 1);
    v->tanfovx = s->tanfovx, v->tanfovy = s->tanfovy;
    v->fx = (float)v->W / (2.0f * s->tanfovx), v->fy = (float)v->H / (2.0f * s->tanfovy);
    v->scale = 1.0f, v->cov_stride = 6, v->sh_ks = 3, v->sh_cs = 1;
    v->view = s->viewmatrix, v->proj = s->projmatrix, v->campos = s->campos, v->bg = s->bg;
    return 0;
}

}  // namespace

extern "C" {

const char* upstream_last_error(void) { return g_msg; }

// Forward on the legacy default stream (as upstream), including the blocking device-to-host read of N.
int upstream_forward(const GgrtRasterSettings* s, int P, const float* means3D, const float* cov3D, const float* opacities,
                     const float* shs, int* radii, float* out_color, float* out_depth, long long* num_rendered) {
    View v;
    make_view(s, P, &v);
    if (int rc = ensure_sizes(P, v.H, v.W)) return rc;
    cudaStream_t st = 0;
    GeomPtrs g = geom_ptrs(g_w.geom, P);
    ImagePtrs im = image_ptrs(g_w.image, v.H, v.W);
    UP_CHECK(cudaMemsetAsync(im.counts, 0, reinterpret_cast<char*>(im.cursor) - reinterpret_cast<char*>(im.counts), st));
    launch_geometry(v, means3D, cov3D, opacities, radii, g, im, st);
    launch_color(v, means3D, shs, nullptr, nullptr, radii, g, st);
    UP_CHECK(cub::DeviceScan::InclusiveSum(g_w.scan_tmp, g_w.scan_bytes, g.tiles, g_w.offsets, P, st));
    uint32_t N = 0;
    UP_CHECK(cudaMemcpy(&N, g_w.offsets + (P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));  // the upstream host sync
    g_w.N = N;
    *num_rendered = N;
    if (int rc = ensure_pairs(N)) return rc;
    UP_CHECK(cudaMemsetAsync(g_w.ranges, 0, sizeof(uint2) * (size_t)g_w.T, st));
    if (N > 0) {
        duplicate_with_keys<<<(P + 255) / 256, 256, 0, st>>>(P, v.gx, g_w.offsets, g.rect, g.tiles, g.rec0, g_w.keys, g_w.vals);
        int bits = 0;
        while ((1 << bits) < g_w.T) ++bits;
        UP_CHECK(cub::DeviceRadixSort::SortPairs(g_w.sort_tmp, g_w.sort_bytes, g_w.keys, g_w.keys_sorted, g_w.vals,
                                                 g_w.vals_sorted, (int)N, 0, 32 + bits, st));
        identify_tile_ranges<<<(N + 255) / 256, 256, 0, st>>>(N, g_w.keys_sorted, g_w.ranges);
    }
    render_forward_classic<<<dim3(v.gx, v.gy), dim3(TILE, TILE), 0, st>>>(v.W, v.H, v.gx, g_w.ranges, g_w.vals_sorted, g.rec0,
                                                                         g.rec1, g.rec2, v.bg, out_color, out_depth,
                                                                         g_w.final_T, g_w.n_contrib);
    UP_CHECK(cudaGetLastError());
    return 0;
}

int upstream_backward(const GgrtRasterSettings* s, int P, const float* means3D, const float* cov3D, const float* shs,
                      const int* radii, const float* dL_dout_color, float* dL_dmeans2D, float* dL_dopacity,
                      float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh) {
    View v;
    make_view(s, P, &v);
    if (P != g_w.P || v.H != g_w.H || v.W != g_w.W) {
        snprintf(g_msg, sizeof(g_msg), "upstream_backward without a matching forward");
        return -1;
    }
    cudaStream_t st = 0;
    GeomPtrs g = geom_ptrs(g_w.geom, P);
    UP_CHECK(cudaMemsetAsync(g_w.scratch, 0, sizeof(float) * (size_t)P * GRAD_STRIDE, st));
    if (g_w.N > 0)
        render_backward_classic<<<dim3(v.gx, v.gy), dim3(TILE, TILE), 0, st>>>(v.W, v.H, v.gx, g_w.ranges, g_w.vals_sorted,
                                                                              g.rec0, g.rec1, g.rec2, v.bg, g_w.final_T,
                                                                              g_w.n_contrib, dL_dout_color, g_w.scratch);
    ColorSinks none;
    memset(&none, 0, sizeof(none));
    launch_preprocess_backward(v, means3D, cov3D, shs, radii, g, g_w.scratch, dL_dmeans2D, dL_dopacity, dL_dmeans3D,
                               dL_dcov3D, dL_dsh, nullptr, nullptr, nullptr, none, st);
    UP_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
