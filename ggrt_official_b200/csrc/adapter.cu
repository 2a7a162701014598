// Fused Gaussian construction: the pixelSplat "Gaussian adapter" that turns the encoder's raw
// per-ray outputs into the rasterizer's inputs, restated as ONE forward and ONE backward kernel.
// Reference: /root/reference/ggrt/model/pixelsplat/encoder/common/gaussian_adapter.py:48-96 (forward),
// gaussians.py:8-44 (quaternion_to_matrix, build_covariance), ggrt/geometry/projection.py:74-114 (unproject,
// get_world_rays), ggrt/misc/sh_rotation.py:10-29 (rotate_sh: per-degree Wigner-D blocks, supplied by the caller).
// The reference runs ~30 small PyTorch kernels with [G,3,3] / [G,3,K] intermediates; here every byte of the
// 348 B/Gaussian output is written once and the raw features are read once.
//
// Indexing ("b v r srf spp" flattened): ray = (view, r, srf) owns `spp` Gaussians g = ray * spp + s that share the
// ray's raw features [7 + 3K] = scales(3) | quaternion xyzw(4) | sh (3 x K, channel-major) and its image
// coordinate; depths are per Gaussian.  A warp takes 8 consecutive rays per iteration: the lanes stream the rays'
// 3K harmonics (coalesced row loads and stores, all loads of the group in flight together; each lane applies the
// <= 9-term Wigner-D block of its coefficient from shared memory), then lane = (ray, sample slot) does the geometry
// of one Gaussian (scale / rotation -> covariance, ray unprojection -> mean).  The backward sums the spp harmonics
// gradients per lane, applies the transposed blocks, and reduces the geometry gradients of a ray's Gaussians over
// its adjacent lanes with warp shuffles -- no atomics.
#include "common.cuh"

namespace ggrt {

constexpr int AD_WARPS = 8;
constexpr int AD_THREADS = 32 * AD_WARPS;
constexpr float AD_QUAT_EPS = 1e-8f;  // quaternion_to_matrix's own eps (gaussians.py:11)

struct AdapterParams {
    int num_views, rays_per_view, spp, image_h, image_w;
    float scale_min, scale_max, eps;
};

struct ViewConsts {
    float C[9];     // camera-to-world rotation (extrinsics[:3,:3])
    float o[3];     // ray origin (extrinsics[:3,3])
    float Ki[9];    // inverse intrinsics
    float mult;     // get_scale_multiplier
};

__device__ __forceinline__ void load_view(const float* __restrict__ E, const float* __restrict__ Kn, int h, int w,
                                          ViewConsts& vc) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) vc.C[3 * i + j] = E[4 * i + j];
        vc.o[i] = E[4 * i + 3];
    }
    // 3x3 inverse by cofactors
    const float a = Kn[0], b = Kn[1], c = Kn[2], d = Kn[3], e = Kn[4], f = Kn[5], g = Kn[6], hh = Kn[7], i = Kn[8];
    const float A = e * i - f * hh, B = -(d * i - f * g), Cc = d * hh - e * g;
    const float det = a * A + b * B + c * Cc, id = 1.0f / det;
    vc.Ki[0] = A * id, vc.Ki[1] = -(b * i - c * hh) * id, vc.Ki[2] = (b * f - c * e) * id;
    vc.Ki[3] = B * id, vc.Ki[4] = (a * i - c * g) * id, vc.Ki[5] = -(a * f - c * d) * id;
    vc.Ki[6] = Cc * id, vc.Ki[7] = -(a * hh - b * g) * id, vc.Ki[8] = (a * e - b * d) * id;
    // multiplier = 0.1 * sum_i (inverse(K[:2,:2]) @ (1/w, 1/h))_i   (gaussian_adapter.py:98-109)
    const float d2 = a * e - b * d, i2 = 1.0f / d2;
    const float px = 1.0f / (float)w, py = 1.0f / (float)h;
    vc.mult = 0.1f * ((e * i2 * px - b * i2 * py) + (-d * i2 * px + a * i2 * py));
}

__device__ __forceinline__ int isqrt_small(int k) { return k >= 16 ? 4 : k >= 9 ? 3 : k >= 4 ? 2 : k >= 1 ? 1 : 0; }
__device__ __forceinline__ float sh_mask_of(int l) {  // gaussian_adapter.py:45-46
    return l == 0 ? 1.0f : l == 1 ? 0.025f : l == 2 ? 0.00625f : l == 3 ? 0.0015625f : 0.000390625f;
}

struct Geo3 {  // forward intermediates of one Gaussian that the backward needs again
    float sg[3], base[3], s[3], R[9], q[4], nq, t, u[3], n, d[3], dw[3];
};

__device__ __forceinline__ void adapter_geometry(const AdapterParams& p, const ViewConsts& vc, const float* raw7, float x,
                                                 float y, float depth, Geo3& g) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        g.sg[j] = 1.0f / (1.0f + expf(-raw7[j]));
        g.base[j] = p.scale_min + (p.scale_max - p.scale_min) * g.sg[j];
        g.s[j] = g.base[j] * depth * vc.mult;
    }
    const float q0 = raw7[3], q1 = raw7[4], q2 = raw7[5], q3 = raw7[6];
    g.nq = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    const float inq = 1.0f / (g.nq + p.eps);
    const float a = q0 * inq, b = q1 * inq, c = q2 * inq, d = q3 * inq;  // (i, j, k, r)
    g.q[0] = a, g.q[1] = b, g.q[2] = c, g.q[3] = d;
    g.t = 2.0f / (a * a + b * b + c * c + d * d + AD_QUAT_EPS);
    const float t = g.t;
    g.R[0] = 1.0f - t * (b * b + c * c), g.R[1] = t * (a * b - c * d), g.R[2] = t * (a * c + b * d);
    g.R[3] = t * (a * b + c * d), g.R[4] = 1.0f - t * (a * a + c * c), g.R[5] = t * (b * c - a * d);
    g.R[6] = t * (a * c - b * d), g.R[7] = t * (b * c + a * d), g.R[8] = 1.0f - t * (a * a + b * b);
#pragma unroll
    for (int i = 0; i < 3; ++i) g.u[i] = vc.Ki[3 * i] * x + vc.Ki[3 * i + 1] * y + vc.Ki[3 * i + 2];
    g.n = sqrtf(g.u[0] * g.u[0] + g.u[1] * g.u[1] + g.u[2] * g.u[2]);
    const float in = 1.0f / g.n;
#pragma unroll
    for (int i = 0; i < 3; ++i) g.d[i] = g.u[i] * in;
#pragma unroll
    for (int i = 0; i < 3; ++i) g.dw[i] = vc.C[3 * i] * g.d[0] + vc.C[3 * i + 1] * g.d[1] + vc.C[3 * i + 2] * g.d[2];
}

// W = C * R * diag(s); world covariance = W W^T
__device__ __forceinline__ void world_factor(const ViewConsts& vc, const Geo3& g, float W[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            W[3 * i + j] = (vc.C[3 * i] * g.R[j] + vc.C[3 * i + 1] * g.R[3 + j] + vc.C[3 * i + 2] * g.R[6 + j]) * g.s[j];
}

// the per-view Wigner-D blocks, compacted: block l starts at sum_{m<l} (2m+1)^2 and is (2l+1) x (2l+1) row-major
__device__ __forceinline__ int block_offset(int l) { return l == 0 ? 0 : l == 1 ? 1 : l == 2 ? 10 : l == 3 ? 35 : 84; }

template <int DEG>
__device__ __forceinline__ void load_blocks(const float* __restrict__ shrot_view, float* sD, int lane) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    constexpr int NB = (DEG + 1) * (2 * DEG + 1) * (2 * DEG + 3) / 3;  // sum of (2l+1)^2
    for (int e = lane; e < NB; e += 32) {
        const int l = e >= 84 ? 4 : e >= 35 ? 3 : e >= 10 ? 2 : e >= 1 ? 1 : 0;
        const int w = 2 * l + 1, r = (e - block_offset(l)) / w, c = (e - block_offset(l)) % w;
        sD[e] = shrot_view ? shrot_view[(l * l + r) * K + l * l + c] : (r == c ? 1.0f : 0.0f);
    }
}

// Rays are processed in groups of AD_RG per warp iteration.  Geometry: lane = (ray of the group, sample slot), so
// with spp = 3 twenty-four lanes work at once instead of three; harmonics: the group's AD_RG rows are fetched
// together (all loads in flight before the first use) and the lanes then stride over the AD_RG x 3K outputs.
constexpr int AD_RG = 8;          // rays per warp iteration
constexpr int AD_SL = 32 / AD_RG;  // sample slots per ray (samples s, s + AD_SL, ... of a ray go to the same lane)

template <int DEG>
__global__ void __launch_bounds__(AD_THREADS, 2)
adapter_forward_kernel(AdapterParams p, const float* __restrict__ extr, const float* __restrict__ intr,
                       const float* __restrict__ shrot, const float* __restrict__ coords,
                       const float* __restrict__ depths, const float* __restrict__ raw, float* __restrict__ means,
                       float* __restrict__ cov, float* __restrict__ harm, float* __restrict__ scales_out,
                       float* __restrict__ rot_out) {
    constexpr int K = (DEG + 1) * (DEG + 1), ROW = 3 * K, CH = 7 + ROW;
    constexpr int NB = (DEG + 1) * (2 * DEG + 1) * (2 * DEG + 3) / 3;
    __shared__ float s_row[AD_WARPS][AD_RG * ROW];
    __shared__ float s_D[AD_WARPS][NB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rays = p.num_views * p.rays_per_view;
    const int groups = (rays + AD_RG - 1) / AD_RG;
    int geo_view = -1, sh_view = -1;  // view whose constants this lane holds / whose blocks the warp holds
    ViewConsts vc;
    for (int grp = blockIdx.x * AD_WARPS + warp; grp < groups; grp += gridDim.x * AD_WARPS) {
        const int ray0 = grp * AD_RG;
        const int nr = min(AD_RG, rays - ray0);
        // ---- harmonics: fetch the group's rows, then mask + per-degree Wigner block, same result for every sample ----
        __syncwarp();  // the previous group's reads of s_row are done
        for (int e = lane; e < nr * ROW; e += 32) {
            const int r = e / ROW, o = e - r * ROW;
            s_row[warp][e] = raw[(size_t)(ray0 + r) * CH + 7 + o];
        }
        __syncwarp();
        for (int r = 0; r < nr; ++r) {
            const int view = (ray0 + r) / p.rays_per_view;
            if (view != sh_view) {  // warp-uniform
                sh_view = view;
                __syncwarp();
                load_blocks<DEG>(shrot ? shrot + (size_t)view * K * K : nullptr, s_D[warp], lane);
                __syncwarp();
            }
            for (int o = lane; o < ROW; o += 32) {
                const int c = o / K, k = o - c * K, l = isqrt_small(k), w = 2 * l + 1;
                const float* blk = s_D[warp] + block_offset(l) + (k - l * l) * w;
                const float* in = s_row[warp] + r * ROW + c * K + l * l;
                float acc = 0.f;
                for (int j = 0; j < w; ++j) acc = fmaf(blk[j], in[j], acc);
                acc *= sh_mask_of(l);
                for (int sI = 0; sI < p.spp; ++sI) harm[((size_t)(ray0 + r) * p.spp + sI) * ROW + o] = acc;
            }
        }
        // ---- geometry: lane = (ray of the group, sample slot) ----
        const int r = lane / AD_SL, ray = ray0 + r;
        if (r < nr) {
            const int view = ray / p.rays_per_view;
            if (view != geo_view) {
                geo_view = view;
                load_view(extr + 16 * view, intr + 9 * view, p.image_h, p.image_w, vc);
            }
            const float* rr = raw + (size_t)ray * CH;
            float r7[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) r7[k] = rr[k];
            const float cx = coords[2 * (size_t)ray], cy = coords[2 * (size_t)ray + 1];
            for (int sI = lane % AD_SL; sI < p.spp; sI += AD_SL) {
                const size_t g = (size_t)ray * p.spp + sI;
                const float depth = depths[g];
                Geo3 q;
                adapter_geometry(p, vc, r7, cx, cy, depth, q);
#pragma unroll
                for (int i = 0; i < 3; ++i) means[3 * g + i] = vc.o[i] + q.dw[i] * depth;
                float W[9];
                world_factor(vc, q, W);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        cov[9 * g + 3 * i + j] = W[3 * i] * W[3 * j] + W[3 * i + 1] * W[3 * j + 1] + W[3 * i + 2] * W[3 * j + 2];
                if (scales_out) scales_out[3 * g] = q.s[0], scales_out[3 * g + 1] = q.s[1], scales_out[3 * g + 2] = q.s[2];
                if (rot_out)
                    rot_out[4 * g] = q.q[0], rot_out[4 * g + 1] = q.q[1], rot_out[4 * g + 2] = q.q[2], rot_out[4 * g + 3] = q.q[3];
            }
        }
    }
}

template <int DEG>
__global__ void __launch_bounds__(AD_THREADS)
adapter_backward_kernel(AdapterParams p, const float* __restrict__ extr, const float* __restrict__ intr,
                        const float* __restrict__ shrot, const float* __restrict__ coords,
                        const float* __restrict__ depths, const float* __restrict__ raw,
                        const float* __restrict__ dmeans, const float* __restrict__ dcov,
                        const float* __restrict__ dharm, float* __restrict__ dcoords, float* __restrict__ ddepths,
                        float* __restrict__ draw) {
    constexpr int K = (DEG + 1) * (DEG + 1), ROW = 3 * K, CH = 7 + ROW;
    constexpr int NB = (DEG + 1) * (2 * DEG + 1) * (2 * DEG + 3) / 3;
    __shared__ float s_row[AD_WARPS][AD_RG * ROW];
    __shared__ float s_D[AD_WARPS][NB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rays = p.num_views * p.rays_per_view;
    const int groups = (rays + AD_RG - 1) / AD_RG;
    int geo_view = -1, sh_view = -1;
    ViewConsts vc;
    for (int grp = blockIdx.x * AD_WARPS + warp; grp < groups; grp += gridDim.x * AD_WARPS) {
        const int ray0 = grp * AD_RG;
        const int nr = min(AD_RG, rays - ray0);
        // ---- harmonics: sum the samples' gradients per ray, then the transposed Wigner block and the mask ----
        __syncwarp();
        for (int e = lane; e < nr * ROW; e += 32) {
            const int r = e / ROW, o = e - r * ROW;
            float acc = 0.f;
            if (dharm)
                for (int sI = 0; sI < p.spp; ++sI) acc += dharm[((size_t)(ray0 + r) * p.spp + sI) * ROW + o];
            s_row[warp][e] = acc;
        }
        __syncwarp();
        for (int r = 0; r < nr; ++r) {
            const int view = (ray0 + r) / p.rays_per_view;
            if (view != sh_view) {
                sh_view = view;
                __syncwarp();
                load_blocks<DEG>(shrot ? shrot + (size_t)view * K * K : nullptr, s_D[warp], lane);
                __syncwarp();
            }
            float* dr = draw + (size_t)(ray0 + r) * CH;
            for (int o = lane; o < ROW; o += 32) {
                const int c = o / K, j = o - c * K, l = isqrt_small(j), w = 2 * l + 1;
                const float* blk = s_D[warp] + block_offset(l) + (j - l * l);  // column j of the block
                const float* in = s_row[warp] + r * ROW + c * K + l * l;
                float acc = 0.f;
                for (int k = 0; k < w; ++k) acc = fmaf(blk[k * w], in[k], acc);
                dr[7 + o] = acc * sh_mask_of(l);
            }
        }
        // ---- geometry: lane = (ray of the group, sample slot); the slots of a ray are adjacent lanes ----
        const int r = lane / AD_SL, ray = ray0 + r;
        float g_raw[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, g_xy[2] = {0.f, 0.f};
        if (r < nr) {
            const int view = ray / p.rays_per_view;
            if (view != geo_view) {
                geo_view = view;
                load_view(extr + 16 * view, intr + 9 * view, p.image_h, p.image_w, vc);
            }
            const float* rr = raw + (size_t)ray * CH;
            float r7[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) r7[k] = rr[k];
            const float cx = coords[2 * (size_t)ray], cy = coords[2 * (size_t)ray + 1];
            for (int sI = lane % AD_SL; sI < p.spp; sI += AD_SL) {
                const size_t g = (size_t)ray * p.spp + sI;
                const float depth = depths[g];
                Geo3 q;
                adapter_geometry(p, vc, r7, cx, cy, depth, q);
                float gdepth = 0.f;
                // mean = o + dw * depth
                float gm[3] = {0.f, 0.f, 0.f};
                if (dmeans) gm[0] = dmeans[3 * g], gm[1] = dmeans[3 * g + 1], gm[2] = dmeans[3 * g + 2];
                gdepth += gm[0] * q.dw[0] + gm[1] * q.dw[1] + gm[2] * q.dw[2];
                float gd[3];  // dL/dd (camera-space unit direction) = C^T (depth * gm)
#pragma unroll
                for (int j = 0; j < 3; ++j) gd[j] = depth * (vc.C[j] * gm[0] + vc.C[3 + j] * gm[1] + vc.C[6 + j] * gm[2]);
                const float dd = q.d[0] * gd[0] + q.d[1] * gd[1] + q.d[2] * gd[2];
                float gu[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) gu[j] = (gd[j] - q.d[j] * dd) / q.n;
                g_xy[0] += vc.Ki[0] * gu[0] + vc.Ki[3] * gu[1] + vc.Ki[6] * gu[2];
                g_xy[1] += vc.Ki[1] * gu[0] + vc.Ki[4] * gu[1] + vc.Ki[7] * gu[2];
                // covariance = W W^T, W = C R diag(s)
                float G[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (dcov)
#pragma unroll
                    for (int k = 0; k < 9; ++k) G[k] = dcov[9 * g + k];
                float W[9], gW[9], gM[9];
                world_factor(vc, q, W);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        float a = 0.f;
#pragma unroll
                        for (int k = 0; k < 3; ++k) a += (G[3 * i + k] + G[3 * k + i]) * W[3 * k + j];
                        gW[3 * i + j] = a;
                    }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        gM[3 * i + j] = vc.C[i] * gW[j] + vc.C[3 + i] * gW[3 + j] + vc.C[6 + i] * gW[6 + j];  // C^T gW
                float gR[9], gs[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    gs[j] = gM[j] * q.R[j] + gM[3 + j] * q.R[3 + j] + gM[6 + j] * q.R[6 + j];
#pragma unroll
                    for (int i = 0; i < 3; ++i) gR[3 * i + j] = gM[3 * i + j] * q.s[j];
                }
                // scales = base * depth * mult, base = smin + (smax - smin) sigmoid(raw)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    gdepth += gs[j] * q.base[j] * vc.mult;
                    g_raw[j] += gs[j] * depth * vc.mult * (p.scale_max - p.scale_min) * q.sg[j] * (1.0f - q.sg[j]);
                }
                // rotation matrix -> normalised quaternion (a, b, c, d) = (i, j, k, r), t = 2 / (|q|^2 + 1e-8)
                const float a = q.q[0], b = q.q[1], c = q.q[2], d = q.q[3], t = q.t;
                const float gt = -gR[0] * (b * b + c * c) + gR[1] * (a * b - c * d) + gR[2] * (a * c + b * d) +
                                 gR[3] * (a * b + c * d) - gR[4] * (a * a + c * c) + gR[5] * (b * c - a * d) +
                                 gR[6] * (a * c - b * d) + gR[7] * (b * c + a * d) - gR[8] * (a * a + b * b);
                float gq[4];
                gq[0] = t * (gR[1] * b + gR[2] * c + gR[3] * b - 2.0f * gR[4] * a - gR[5] * d + gR[6] * c + gR[7] * d - 2.0f * gR[8] * a);
                gq[1] = t * (-2.0f * gR[0] * b + gR[1] * a + gR[2] * d + gR[3] * a + gR[5] * c - gR[6] * d + gR[7] * c - 2.0f * gR[8] * b);
                gq[2] = t * (-2.0f * gR[0] * c - gR[1] * d + gR[2] * a + gR[3] * d - 2.0f * gR[4] * c + gR[5] * b + gR[6] * a + gR[7] * b);
                gq[3] = t * (-gR[1] * c + gR[2] * b + gR[3] * c - gR[5] * a - gR[6] * b + gR[7] * a);
                const float tt = gt * (-t * t);
#pragma unroll
                for (int k = 0; k < 4; ++k) gq[k] += tt * q.q[k];
                // q = raw / (|raw| + eps)
                const float den = q.nq + p.eps;
                const float dot = gq[0] * r7[3] + gq[1] * r7[4] + gq[2] * r7[5] + gq[3] * r7[6];
                const float corr = q.nq > 0.f ? dot / (q.nq * den * den) : 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) g_raw[3 + k] += gq[k] / den - r7[3 + k] * corr;
                ddepths[g] = gdepth;
            }
        }
        // sum over the AD_SL sample slots of each ray (adjacent lanes); every lane takes part in the shuffles
#pragma unroll
        for (int off = 1; off < AD_SL; off <<= 1) {
#pragma unroll
            for (int k = 0; k < 7; ++k) g_raw[k] += __shfl_xor_sync(0xffffffffu, g_raw[k], off);
            g_xy[0] += __shfl_xor_sync(0xffffffffu, g_xy[0], off);
            g_xy[1] += __shfl_xor_sync(0xffffffffu, g_xy[1], off);
        }
        if (r < nr && lane % AD_SL == 0) {
            float* dr = draw + (size_t)ray * CH;
#pragma unroll
            for (int k = 0; k < 7; ++k) dr[k] = g_raw[k];
            dcoords[2 * (size_t)ray] = g_xy[0];
            dcoords[2 * (size_t)ray + 1] = g_xy[1];
        }
    }
}

static int adapter_grid(int rays) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = ((rays + AD_RG - 1) / AD_RG + AD_WARPS - 1) / AD_WARPS;
    return want < 8 * sms ? (want > 0 ? want : 1) : 8 * sms;
}

void launch_adapter_forward(const GgrtAdapterParams& hp, const float* extr, const float* intr, const float* shrot,
                            const float* coords, const float* depths, const float* raw, float* means, float* cov,
                            float* harm, float* scales_out, float* rot_out, cudaStream_t s) {
    AdapterParams p{hp.num_views, hp.rays_per_view, hp.samples_per_ray, hp.image_height, hp.image_width,
                    hp.scale_min, hp.scale_max, hp.eps};
    const int rays = p.num_views * p.rays_per_view;
    if (rays == 0 || p.spp == 0) return;
    const int grid = adapter_grid(rays);
#define GGRT_AD_FWD(D) adapter_forward_kernel<D><<<grid, AD_THREADS, 0, s>>>(p, extr, intr, shrot, coords, depths, raw, means, cov, harm, scales_out, rot_out)
    switch (hp.sh_degree) {
        case 0: GGRT_AD_FWD(0); break;
        case 1: GGRT_AD_FWD(1); break;
        case 2: GGRT_AD_FWD(2); break;
        case 3: GGRT_AD_FWD(3); break;
        default: GGRT_AD_FWD(4); break;
    }
#undef GGRT_AD_FWD
}

void launch_adapter_backward(const GgrtAdapterParams& hp, const float* extr, const float* intr, const float* shrot,
                             const float* coords, const float* depths, const float* raw, const float* dmeans,
                             const float* dcov, const float* dharm, float* dcoords, float* ddepths, float* draw,
                             cudaStream_t s) {
    AdapterParams p{hp.num_views, hp.rays_per_view, hp.samples_per_ray, hp.image_height, hp.image_width,
                    hp.scale_min, hp.scale_max, hp.eps};
    const int rays = p.num_views * p.rays_per_view;
    if (rays == 0 || p.spp == 0) return;
    const int grid = adapter_grid(rays);
#define GGRT_AD_BWD(D) adapter_backward_kernel<D><<<grid, AD_THREADS, 0, s>>>(p, extr, intr, shrot, coords, depths, raw, dmeans, dcov, dharm, dcoords, ddepths, draw)
    switch (hp.sh_degree) {
        case 0: GGRT_AD_BWD(0); break;
        case 1: GGRT_AD_BWD(1); break;
        case 2: GGRT_AD_BWD(2); break;
        case 3: GGRT_AD_BWD(3); break;
        default: GGRT_AD_BWD(4); break;
    }
#undef GGRT_AD_BWD
}

}  // namespace ggrt
