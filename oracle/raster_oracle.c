/*
 * raster_oracle.c -- CPU restatement of the differentiable 3D-Gaussian tile rasterizer
 * that GGRt calls at ggrt/model/pixelsplat/decoder/cuda_splatting.py:101-125.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (ggrt_official_b200/) never imports, links or executes anything under oracle/.
 *
 * PARITY UNPINNED: the algorithm lives in an un-vendored, un-pinned third-party CUDA
 * package (dcharatan/diff-gaussian-rasterization-modified, README.md:17-18 of the
 * reference) whose source is not under /root/reference, and the reference ships no
 * tests or golden vectors for it.  This file restates the
 * published 3DGS tile-rasterizer algorithm (Kerbl et al. 2023) as specified in
 * SURVEY.md Appendix A.1-A.5; it is cross-checked against an independent autograd
 * restatement (oracle/torch_ref.py) and closed-form cases (tests/test_oracle_*.py).
 *
 * Conventions: row-vector maths, viewmatrix V and projmatrix M
 * are the 4x4 tensors GGRt passes (cuda_splatting.py:108-109), row-major, translation in
 * the last row: p_view = [p,1].V.  cov3D is (xx,xy,xz,yy,yz,zz) (cuda_splatting.py:116,124).
 *
 * Geometry arithmetic (cull, cov2D, radius, tile rect, depth key) uses only IEEE
 * float add/sub/mul/div/sqrt in a fixed order and NO fused multiply-add (compile with
 * -ffp-contract=off), so that a device implementation using round-to-nearest
 * intrinsics in the same order is bit-identical: radii, tile rects, sort keys and the
 * sorted per-tile lists are compared bit-exactly.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define NEAR_CULL 0.2f        /* stock 3DGS in_frustum threshold (A.1) */
#define LOWPASS 0.3f          /* screen-space dilation added to cov2D diagonal (A.1) */
#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_MAX 0.99f
#define T_EPS 0.0001f
#define RADIUS_CAP 1.0e6f     /* guards the float->int cast for absurd covariances */

/* relative band around a hard threshold inside which a pixel is flagged "fragile":
 * an implementation whose exp()/fma rounding differs by a few ulp may legitimately
 * take the other branch there (A.3 thresholds are discontinuities of the algorithm). */
#define FRAGILE_REL 2.0e-5f

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};
static const float SH_C4[9] = {2.5033429417967046f, -1.7701307697799304f, 0.9461746957575601f,
                               -0.6690465435572892f, 0.10578554691520431f, -0.6690465435572892f,
                               0.47308734787878004f, -1.7701307697799304f, 0.6258357354491761f};

typedef struct {
    int P;            /* number of Gaussians */
    int deg;          /* SH degree 0..4 ; K = (deg+1)^2 coefficients per channel */
    int W, H;
    float tanfovx, tanfovy;
    float view[16];
    float proj[16];
    float campos[3];
    float bg[3];
} OracleCam;

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* float -> int with the saturation made explicit (C's cast is UB out of range) */
static inline int f2i_sat(float v) {
    v = fmaxf(v, -1.0f);
    v = fminf(v, 65536.0f);
    return (int)v;
}

/* ------------------------------------------------------------------------------------
 * A.1 preprocess: per Gaussian cull / project / cov2D / conic / radius / rect / colour.
 * Outputs (all preallocated, zeroed here):
 *   depth[P], radii[P] (int), xy[P*2], conic_opacity[P*4], rgb[P*3], clamped[P*3] (u8),
 *   rect[P*4] (int: xmin,ymin,xmax,ymax), tiles_touched[P] (u32)
 * sh is [P,K,3] (cuda_splatting.py:77) or NULL when colors_precomp [P,3] is given.
 * ---------------------------------------------------------------------------------- */
static void eval_sh(int deg, const float* sh /* K*3 */, float x, float y, float z, float out[3]) {
    for (int c = 0; c < 3; ++c) {
        float r = SH_C0 * sh[0 * 3 + c];
        if (deg > 0) {
            r = r - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
                    SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
                    SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] +
                        SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] + SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
                    if (deg > 3) {
                        r = r + SH_C4[0] * xy * (xx - yy) * sh[16 * 3 + c] +
                            SH_C4[1] * yz * (3.0f * xx - yy) * sh[17 * 3 + c] +
                            SH_C4[2] * xy * (7.0f * zz - 1.0f) * sh[18 * 3 + c] +
                            SH_C4[3] * yz * (7.0f * zz - 3.0f) * sh[19 * 3 + c] +
                            SH_C4[4] * (zz * (35.0f * zz - 30.0f) + 3.0f) * sh[20 * 3 + c] +
                            SH_C4[5] * xz * (7.0f * zz - 3.0f) * sh[21 * 3 + c] +
                            SH_C4[6] * (xx - yy) * (7.0f * zz - 1.0f) * sh[22 * 3 + c] +
                            SH_C4[7] * xz * (xx - 3.0f * yy) * sh[23 * 3 + c] +
                            SH_C4[8] * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy)) * sh[24 * 3 + c];
                    }
                }
            }
        }
        out[c] = r;
    }
}

/* the shared geometric core; returns 0 if culled. No FMA, fixed op order. */
typedef struct {
    float tx, ty, tz;          /* camera-space mean (unclamped) */
    float cx, cy;              /* frustum-clamped tx, ty (A.1) */
    float txtz, tytz;
    float hx, hy, hw, pw;      /* homogeneous clip coords and 1/(w+1e-7) */
    float Tm[2][3];            /* J.Rw */
    float a, b, c, det;        /* dilated cov2D */
    float fx, fy;
} Geo;

static int geometry(const OracleCam* cam, const float* p, const float* cv, Geo* g) {
    const float* V = cam->view;
    const float* M = cam->proj;
    float px = p[0], py = p[1], pz = p[2];
    g->tx = ((V[0] * px + V[4] * py) + V[8] * pz) + V[12];
    g->ty = ((V[1] * px + V[5] * py) + V[9] * pz) + V[13];
    g->tz = ((V[2] * px + V[6] * py) + V[10] * pz) + V[14];
    if (!(g->tz > NEAR_CULL)) return 0;
    g->hx = ((M[0] * px + M[4] * py) + M[8] * pz) + M[12];
    g->hy = ((M[1] * px + M[5] * py) + M[9] * pz) + M[13];
    g->hw = ((M[3] * px + M[7] * py) + M[11] * pz) + M[15];
    g->pw = 1.0f / (g->hw + 0.0000001f);
    g->fx = (float)cam->W / (2.0f * cam->tanfovx);
    g->fy = (float)cam->H / (2.0f * cam->tanfovy);
    float limx = 1.3f * cam->tanfovx, limy = 1.3f * cam->tanfovy;
    g->txtz = g->tx / g->tz;
    g->tytz = g->ty / g->tz;
    g->cx = fminf(limx, fmaxf(-limx, g->txtz)) * g->tz;
    g->cy = fminf(limy, fmaxf(-limy, g->tytz)) * g->tz;
    float tz2 = g->tz * g->tz;
    float J00 = g->fx / g->tz, J02 = -(g->fx * g->cx) / tz2;
    float J11 = g->fy / g->tz, J12 = -(g->fy * g->cy) / tz2;
    for (int k = 0; k < 3; ++k) {
        float r0 = V[4 * k + 0], r1 = V[4 * k + 1], r2 = V[4 * k + 2];
        g->Tm[0][k] = J00 * r0 + J02 * r2;
        g->Tm[1][k] = J11 * r1 + J12 * r2;
    }
    float s00 = cv[0], s01 = cv[1], s02 = cv[2], s11 = cv[3], s12 = cv[4], s22 = cv[5];
    const float* t0 = g->Tm[0];
    const float* t1 = g->Tm[1];
    float v00 = (s00 * t0[0] + s01 * t0[1]) + s02 * t0[2];
    float v01 = (s01 * t0[0] + s11 * t0[1]) + s12 * t0[2];
    float v02 = (s02 * t0[0] + s12 * t0[1]) + s22 * t0[2];
    float v10 = (s00 * t1[0] + s01 * t1[1]) + s02 * t1[2];
    float v11 = (s01 * t1[0] + s11 * t1[1]) + s12 * t1[2];
    float v12 = (s02 * t1[0] + s12 * t1[1]) + s22 * t1[2];
    g->a = ((t0[0] * v00 + t0[1] * v01) + t0[2] * v02) + LOWPASS;
    g->b = (t1[0] * v00 + t1[1] * v01) + t1[2] * v02;
    g->c = ((t1[0] * v10 + t1[1] * v11) + t1[2] * v12) + LOWPASS;
    g->det = g->a * g->c - g->b * g->b;
    if (!(g->det > 0.0f) && !(g->det < 0.0f)) return 0; /* det == 0 (A.1) or NaN */
    return 1;
}

void oracle_preprocess(const OracleCam* cam, const float* means, const float* cov3d, const float* opac,
                       const float* sh, const float* colors_precomp, float* depth, int* radii, float* xy,
                       float* conic_opacity, float* rgb, uint8_t* clamped, int* rect, uint32_t* tiles_touched) {
    const int P = cam->P;
    const int K = (cam->deg + 1) * (cam->deg + 1);
    const int gx = (cam->W + TILE - 1) / TILE, gy = (cam->H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        depth[i] = 0.0f;
        radii[i] = 0;
        tiles_touched[i] = 0;
        xy[2 * i] = xy[2 * i + 1] = 0.0f;
        for (int k = 0; k < 4; ++k) conic_opacity[4 * i + k] = 0.0f, rect[4 * i + k] = 0;
        for (int k = 0; k < 3; ++k) rgb[3 * i + k] = 0.0f, clamped[3 * i + k] = 0;
        Geo g;
        if (!geometry(cam, means + 3 * i, cov3d + 6 * i, &g)) continue;
        float dinv = 1.0f / g.det;
        float cA = g.c * dinv, cB = -g.b * dinv, cC = g.a * dinv;
        float mid = 0.5f * (g.a + g.c);
        float sq = sqrtf(fmaxf(0.1f, mid * mid - g.det));
        float l1 = mid + sq, l2 = mid - sq;
        float radf = fminf(ceilf(3.0f * sqrtf(fmaxf(l1, l2))), RADIUS_CAP);
        float ndcx = g.hx * g.pw, ndcy = g.hy * g.pw;
        float pxx = ((ndcx + 1.0f) * (float)cam->W - 1.0f) * 0.5f;
        float pxy = ((ndcy + 1.0f) * (float)cam->H - 1.0f) * 0.5f;
        int x0 = clampi(f2i_sat((pxx - radf) / (float)TILE), 0, gx);
        int y0 = clampi(f2i_sat((pxy - radf) / (float)TILE), 0, gy);
        int x1 = clampi(f2i_sat((pxx + radf + (float)(TILE - 1)) / (float)TILE), 0, gx);
        int y1 = clampi(f2i_sat((pxy + radf + (float)(TILE - 1)) / (float)TILE), 0, gy);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (colors_precomp) {
            for (int c = 0; c < 3; ++c) rgb[3 * i + c] = colors_precomp[3 * i + c];
        } else {
            float dx = means[3 * i] - cam->campos[0], dy = means[3 * i + 1] - cam->campos[1],
                  dz = means[3 * i + 2] - cam->campos[2];
            float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            float col[3];
            eval_sh(cam->deg, sh + (size_t)i * K * 3, dx * inv, dy * inv, dz * inv, col);
            for (int c = 0; c < 3; ++c) {
                float v = col[c] + 0.5f;
                clamped[3 * i + c] = (uint8_t)(v < 0.0f);
                rgb[3 * i + c] = fmaxf(v, 0.0f);
            }
        }
        depth[i] = g.tz;
        radii[i] = (int)radf;
        xy[2 * i] = pxx;
        xy[2 * i + 1] = pxy;
        conic_opacity[4 * i + 0] = cA;
        conic_opacity[4 * i + 1] = cB;
        conic_opacity[4 * i + 2] = cC;
        conic_opacity[4 * i + 3] = opac[i];
        rect[4 * i + 0] = x0;
        rect[4 * i + 1] = y0;
        rect[4 * i + 2] = x1;
        rect[4 * i + 3] = y1;
        tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
    }
}

/* ------------------------------------------------------------------------------------
 * A.2 binning: inclusive scan, key emit (y outer, x inner), stable LSD radix sort,
 * per-tile [start,end) ranges.
 * ---------------------------------------------------------------------------------- */
uint64_t oracle_count_pairs(int P, const uint32_t* tiles_touched, uint32_t* offsets /* inclusive scan, [P] */) {
    uint64_t run = 0;
    for (int i = 0; i < P; ++i) {
        run += tiles_touched[i];
        offsets[i] = (uint32_t)run;
    }
    return run;
}

static void radix_sort_pairs(uint64_t* keys, uint32_t* vals, size_t n) {
    uint64_t* k2 = (uint64_t*)malloc(n * sizeof(uint64_t));
    uint32_t* v2 = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint64_t *ka = keys, *kb = k2;
    uint32_t *va = vals, *vb = v2;
    for (int pass = 0; pass < 8; ++pass) {
        size_t hist[257];
        memset(hist, 0, sizeof(hist));
        int sh = 8 * pass;
        for (size_t i = 0; i < n; ++i) hist[((ka[i] >> sh) & 0xff) + 1]++;
        for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
        for (size_t i = 0; i < n; ++i) {
            size_t d = hist[(ka[i] >> sh) & 0xff]++;
            kb[d] = ka[i];
            vb[d] = va[i];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    /* 8 passes: result is back in the caller's arrays */
    free(k2);
    free(v2);
}

void oracle_bin(int P, int W, int H, const float* depth, const int* radii, const int* rect, const uint32_t* offsets,
                uint64_t N, uint64_t* keys /* [N] out, sorted */, uint32_t* vals /* [N] out, sorted */,
                uint32_t* ranges /* [T*2] out */) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    for (int i = 0; i < P; ++i) {
        if (radii[i] <= 0) continue;
        uint64_t off = (i == 0) ? 0 : offsets[i - 1];
        uint32_t dbits;
        memcpy(&dbits, &depth[i], 4);
        for (int y = rect[4 * i + 1]; y < rect[4 * i + 3]; ++y)
            for (int x = rect[4 * i + 0]; x < rect[4 * i + 2]; ++x) {
                uint64_t key = (uint64_t)(y * gx + x);
                keys[off] = (key << 32) | dbits;
                vals[off] = (uint32_t)i;
                ++off;
            }
    }
    radix_sort_pairs(keys, vals, (size_t)N);
    memset(ranges, 0, (size_t)gx * gy * 2 * sizeof(uint32_t));
    for (uint64_t j = 0; j < N; ++j) {
        uint32_t t = (uint32_t)(keys[j] >> 32);
        if (j == 0 || (uint32_t)(keys[j - 1] >> 32) != t) ranges[2 * t] = (uint32_t)j;
        if (j + 1 == N || (uint32_t)(keys[j + 1] >> 32) != t) ranges[2 * t + 1] = (uint32_t)(j + 1);
    }
}

/* ------------------------------------------------------------------------------------
 * A.3 render forward.  out_color[3,H,W], out_depth[H,W], final_T[H,W], n_contrib[H,W];
 * fragile[H,W] (u8, may be NULL) is set where any threshold test of that pixel fell
 * within FRAGILE_REL of its boundary.
 * ---------------------------------------------------------------------------------- */
static inline int near_rel(float v, float thr) { return fabsf(v - thr) <= FRAGILE_REL * fabsf(thr); }

void oracle_render_forward(const OracleCam* cam, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                           const float* conic_opacity, const float* rgb, const float* depth, float* out_color,
                           float* out_depth, float* final_T, uint32_t* n_contrib, uint8_t* fragile,
                           uint8_t* fragile_gaussian) {
    /* fragile_gaussian[P] (u8, may be NULL, zeroed by the caller): bit 0 is set for the Gaussian whose own threshold
     * test was within the band at some pixel -- the one whose contribution an implementation with another rounding
     * may gain or lose there -- and bit 1 for every Gaussian that contributes to such a pixel: if the flagged one
     * flips, the transmittance in front of the ones behind it and the colour accumulated behind the ones in front of
     * it change by up to alpha ~ 0.4 %, which is more than the 1e-3 gradient bar for a splat that covers only a
     * pixel or two.  Several threads may OR the same bits concurrently; that is benign. */
    const int W = cam->W, H = cam->H;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int ty = tile / gx, tx = tile % gx;
        uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                float pxf = (float)x, pyf = (float)y;
                float T = 1.0f, C[3] = {0, 0, 0}, D = 0.0f;
                uint32_t contributor = 0, last = 0;
                uint8_t frag = 0, frag_alpha = 0;
                for (uint32_t j = r0; j < r1; ++j) {
                    ++contributor;
                    uint32_t g = point_list[j];
                    float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                    float cA = conic_opacity[4 * g], cB = conic_opacity[4 * g + 1], cC = conic_opacity[4 * g + 2],
                          o = conic_opacity[4 * g + 3];
                    float power = -0.5f * (cA * dx * dx + cC * dy * dy) - cB * dx * dy;
                    uint8_t fg = 0;
                    if (fabsf(power) <= 1.0e-6f) fg = 1;
                    float oG = power > 0.0f ? 0.0f : o * expf(power);
                    if (power <= 0.0f && (near_rel(oG, ALPHA_MAX) || near_rel(oG, ALPHA_MIN))) fg = 1;
                    if (fg) {
                        frag = 1;
                        frag_alpha = 1;
                        if (fragile_gaussian) fragile_gaussian[g] |= 1;
                    }
                    if (power > 0.0f) continue;
                    float alpha = fminf(ALPHA_MAX, oG);
                    if (alpha < ALPHA_MIN) continue;
                    float Tn = T * (1.0f - alpha);
                    if (near_rel(Tn, T_EPS)) {
                        frag = 1;
                        if (fragile_gaussian) fragile_gaussian[g] |= 1;
                    }
                    if (Tn < T_EPS) break;
                    float w = alpha * T;
                    for (int c = 0; c < 3; ++c) C[c] += rgb[3 * g + c] * w;
                    D += depth[g] * w;
                    T = Tn;
                    last = contributor;
                }
                size_t pix = (size_t)y * W + x;
                for (int c = 0; c < 3; ++c) out_color[(size_t)c * H * W + pix] = C[c] + T * cam->bg[c];
                out_depth[pix] = D;
                final_T[pix] = T;
                n_contrib[pix] = last;
                if (fragile) fragile[pix] = frag;
                if (frag_alpha && fragile_gaussian) { /* flag every contributor of this pixel (bit 1) */
                    for (uint32_t j = r0; j < r1 && j < r0 + last + 1; ++j) {
                        uint32_t g = point_list[j];
                        float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                        float power = -0.5f * (conic_opacity[4 * g] * dx * dx + conic_opacity[4 * g + 2] * dy * dy) -
                                      conic_opacity[4 * g + 1] * dx * dy;
                        if (power > 0.0f) continue;
                        if (conic_opacity[4 * g + 3] * expf(power) >= ALPHA_MIN * (1.0f - FRAGILE_REL))
                            fragile_gaussian[g] |= 2;
                    }
                }
            }
    }
}

/* ------------------------------------------------------------------------------------
 * A.4 render backward.  dL_dout_color[3,H,W]; dL_dout_depth[H,W] may be NULL.
 * Accumulates (zeroing first) dL_dmean2D[P*2] (w.r.t. NDC xy), dL_dconic[P*3] (A, B/2
 * convention, C), dL_dopacity[P], dL_dcolor[P*3], dL_ddepth[P] (may be NULL iff
 * dL_dout_depth is NULL).  Per-pixel maths in float, cross-pixel sums in double.
 * ---------------------------------------------------------------------------------- */
void oracle_render_backward(const OracleCam* cam, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                            const float* conic_opacity, const float* rgb, const float* depth, const float* final_T,
                            const uint32_t* n_contrib, const float* dL_dout_color, const float* dL_dout_depth,
                            float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                            float* dL_ddepth) {
    const int W = cam->W, H = cam->H, P = cam->P;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    enum { NV = 10 }; /* mean2D.xy, conic.ABC, opacity, color.rgb, depth */
    double* acc = (double*)calloc((size_t)P * NV, sizeof(double));
#pragma omp parallel
    {
        double* loc = NULL;
        size_t loc_cap = 0;
#pragma omp for schedule(dynamic, 4)
        for (int tile = 0; tile < gx * gy; ++tile) {
            int ty = tile / gx, tx = tile % gx;
            uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
            size_t n = r1 - r0;
            if (n == 0) continue;
            if (n > loc_cap) {
                free(loc);
                loc = (double*)malloc(n * NV * sizeof(double));
                loc_cap = n;
            }
            memset(loc, 0, n * NV * sizeof(double));
            for (int ly = 0; ly < TILE; ++ly)
                for (int lx = 0; lx < TILE; ++lx) {
                    int x = tx * TILE + lx, y = ty * TILE + ly;
                    if (x >= W || y >= H) continue;
                    size_t pix = (size_t)y * W + x;
                    float pxf = (float)x, pyf = (float)y;
                    float T_final = final_T[pix];
                    float T = T_final;
                    uint32_t last = n_contrib[pix];
                    float dpix[3], ddep = dL_dout_depth ? dL_dout_depth[pix] : 0.0f;
                    for (int c = 0; c < 3; ++c) dpix[c] = dL_dout_color[(size_t)c * H * W + pix];
                    float bg_dot = cam->bg[0] * dpix[0] + cam->bg[1] * dpix[1] + cam->bg[2] * dpix[2];
                    float accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0};
                    float accum_d = 0.0f, last_d = 0.0f, last_alpha = 0.0f;
                    for (uint32_t jj = last; jj-- > 0;) {
                        uint32_t g = point_list[r0 + jj];
                        float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                        float cA = conic_opacity[4 * g], cB = conic_opacity[4 * g + 1],
                              cC = conic_opacity[4 * g + 2], o = conic_opacity[4 * g + 3];
                        float power = -0.5f * (cA * dx * dx + cC * dy * dy) - cB * dx * dy;
                        if (power > 0.0f) continue;
                        float G = expf(power);
                        float alpha = fminf(ALPHA_MAX, o * G);
                        if (alpha < ALPHA_MIN) continue;
                        T = T / (1.0f - alpha);
                        float w = alpha * T;
                        float dL_dalpha = 0.0f;
                        double* L = loc + (size_t)jj * NV;
                        for (int c = 0; c < 3; ++c) {
                            float col = rgb[3 * g + c];
                            accum[c] = last_alpha * last_color[c] + (1.0f - last_alpha) * accum[c];
                            last_color[c] = col;
                            dL_dalpha += (col - accum[c]) * dpix[c];
                            L[6 + c] += (double)(w * dpix[c]);
                        }
                        if (dL_dout_depth) {
                            float dep = depth[g];
                            accum_d = last_alpha * last_d + (1.0f - last_alpha) * accum_d;
                            last_d = dep;
                            dL_dalpha += (dep - accum_d) * ddep;
                            L[9] += (double)(w * ddep);
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                        float dL_dG = o * dL_dalpha;
                        float gdx = G * dx, gdy = G * dy;
                        float dG_ddelx = -gdx * cA - gdy * cB;
                        float dG_ddely = -gdy * cC - gdx * cB;
                        L[0] += (double)(dL_dG * dG_ddelx * (0.5f * (float)W));
                        L[1] += (double)(dL_dG * dG_ddely * (0.5f * (float)H));
                        L[2] += (double)(-0.5f * gdx * dx * dL_dG);
                        L[3] += (double)(-0.5f * gdx * dy * dL_dG);
                        L[4] += (double)(-0.5f * gdy * dy * dL_dG);
                        L[5] += (double)(G * dL_dalpha);
                    }
                }
            for (size_t j = 0; j < n; ++j) {
                uint32_t g = point_list[r0 + j];
                for (int k = 0; k < NV; ++k) {
                    double v = loc[j * NV + k];
                    if (v != 0.0) {
#pragma omp atomic
                        acc[(size_t)g * NV + k] += v;
                    }
                }
            }
        }
        free(loc);
    }
    for (int i = 0; i < P; ++i) {
        const double* a = acc + (size_t)i * NV;
        dL_dmean2D[2 * i] = (float)a[0];
        dL_dmean2D[2 * i + 1] = (float)a[1];
        dL_dconic[3 * i] = (float)a[2];
        dL_dconic[3 * i + 1] = (float)a[3];
        dL_dconic[3 * i + 2] = (float)a[4];
        dL_dopacity[i] = (float)a[5];
        dL_dcolor[3 * i] = (float)a[6];
        dL_dcolor[3 * i + 1] = (float)a[7];
        dL_dcolor[3 * i + 2] = (float)a[8];
        if (dL_ddepth) dL_ddepth[i] = (float)a[9];
    }
    free(acc);
}

/* ------------------------------------------------------------------------------------
 * A.5 per-Gaussian backward: conic -> cov2D -> cov3D and mean (through J); NDC mean ->
 * mean3D through M; colour -> SH coefficients and mean3D (through the view direction).
 * Outputs (overwritten): dL_dmeans[P*3], dL_dcov3d[P*6], dL_dsh[P*K*3] (NULL when
 * colours were precomputed).  dL_ddepth (per-Gaussian, may be NULL) flows to the mean
 * through t.z.
 * ---------------------------------------------------------------------------------- */
static void sh_backward(int deg, const float* sh, float x, float y, float z, const float dL_dRGB[3], float* dL_dsh,
                        float dL_ddir[3]) {
    /* basis values b[k] and their partial derivatives w.r.t. the (free) direction components */
    float b[25], bx[25], by[25], bz[25];
    int K = (deg + 1) * (deg + 1);
    for (int k = 0; k < 25; ++k) b[k] = bx[k] = by[k] = bz[k] = 0.0f;
    b[0] = SH_C0;
    if (deg > 0) {
        b[1] = -SH_C1 * y; by[1] = -SH_C1;
        b[2] = SH_C1 * z;  bz[2] = SH_C1;
        b[3] = -SH_C1 * x; bx[3] = -SH_C1;
    }
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    if (deg > 1) {
        b[4] = SH_C2[0] * xy;                     bx[4] = SH_C2[0] * y;  by[4] = SH_C2[0] * x;
        b[5] = SH_C2[1] * yz;                     by[5] = SH_C2[1] * z;  bz[5] = SH_C2[1] * y;
        b[6] = SH_C2[2] * (2.0f * zz - xx - yy);  bx[6] = SH_C2[2] * -2.0f * x; by[6] = SH_C2[2] * -2.0f * y; bz[6] = SH_C2[2] * 4.0f * z;
        b[7] = SH_C2[3] * xz;                     bx[7] = SH_C2[3] * z;  bz[7] = SH_C2[3] * x;
        b[8] = SH_C2[4] * (xx - yy);              bx[8] = SH_C2[4] * 2.0f * x; by[8] = SH_C2[4] * -2.0f * y;
    }
    if (deg > 2) {
        b[9] = SH_C3[0] * y * (3.0f * xx - yy);   bx[9] = SH_C3[0] * 6.0f * xy; by[9] = SH_C3[0] * (3.0f * xx - 3.0f * yy);
        b[10] = SH_C3[1] * xy * z;                bx[10] = SH_C3[1] * yz; by[10] = SH_C3[1] * xz; bz[10] = SH_C3[1] * xy;
        b[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
        bx[11] = SH_C3[2] * -2.0f * xy; by[11] = SH_C3[2] * (4.0f * zz - xx - 3.0f * yy); bz[11] = SH_C3[2] * 8.0f * yz;
        b[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        bx[12] = SH_C3[3] * -6.0f * xz; by[12] = SH_C3[3] * -6.0f * yz; bz[12] = SH_C3[3] * (6.0f * zz - 3.0f * xx - 3.0f * yy);
        b[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
        bx[13] = SH_C3[4] * (4.0f * zz - 3.0f * xx - yy); by[13] = SH_C3[4] * -2.0f * xy; bz[13] = SH_C3[4] * 8.0f * xz;
        b[14] = SH_C3[5] * z * (xx - yy);         bx[14] = SH_C3[5] * 2.0f * xz; by[14] = SH_C3[5] * -2.0f * yz; bz[14] = SH_C3[5] * (xx - yy);
        b[15] = SH_C3[6] * x * (xx - 3.0f * yy);  bx[15] = SH_C3[6] * (3.0f * xx - 3.0f * yy); by[15] = SH_C3[6] * -6.0f * xy;
    }
    if (deg > 3) {
        b[16] = SH_C4[0] * xy * (xx - yy);
        bx[16] = SH_C4[0] * y * (3.0f * xx - yy); by[16] = SH_C4[0] * x * (xx - 3.0f * yy);
        b[17] = SH_C4[1] * yz * (3.0f * xx - yy);
        bx[17] = SH_C4[1] * 6.0f * xy * z; by[17] = SH_C4[1] * z * (3.0f * xx - 3.0f * yy); bz[17] = SH_C4[1] * y * (3.0f * xx - yy);
        b[18] = SH_C4[2] * xy * (7.0f * zz - 1.0f);
        bx[18] = SH_C4[2] * y * (7.0f * zz - 1.0f); by[18] = SH_C4[2] * x * (7.0f * zz - 1.0f); bz[18] = SH_C4[2] * 14.0f * xy * z;
        b[19] = SH_C4[3] * yz * (7.0f * zz - 3.0f);
        by[19] = SH_C4[3] * z * (7.0f * zz - 3.0f); bz[19] = SH_C4[3] * y * (21.0f * zz - 3.0f);
        b[20] = SH_C4[4] * (zz * (35.0f * zz - 30.0f) + 3.0f);
        bz[20] = SH_C4[4] * (140.0f * zz * z - 60.0f * z);
        b[21] = SH_C4[5] * xz * (7.0f * zz - 3.0f);
        bx[21] = SH_C4[5] * z * (7.0f * zz - 3.0f); bz[21] = SH_C4[5] * x * (21.0f * zz - 3.0f);
        b[22] = SH_C4[6] * (xx - yy) * (7.0f * zz - 1.0f);
        bx[22] = SH_C4[6] * 2.0f * x * (7.0f * zz - 1.0f); by[22] = SH_C4[6] * -2.0f * y * (7.0f * zz - 1.0f);
        bz[22] = SH_C4[6] * (xx - yy) * 14.0f * z;
        b[23] = SH_C4[7] * xz * (xx - 3.0f * yy);
        bx[23] = SH_C4[7] * z * (3.0f * xx - 3.0f * yy); by[23] = SH_C4[7] * -6.0f * xy * z; bz[23] = SH_C4[7] * x * (xx - 3.0f * yy);
        b[24] = SH_C4[8] * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
        bx[24] = SH_C4[8] * (4.0f * xx * x - 12.0f * x * yy); by[24] = SH_C4[8] * (-12.0f * xx * y + 4.0f * yy * y);
    }
    dL_ddir[0] = dL_ddir[1] = dL_ddir[2] = 0.0f;
    for (int k = 0; k < K; ++k)
        for (int c = 0; c < 3; ++c) {
            dL_dsh[k * 3 + c] = b[k] * dL_dRGB[c];
            float s = sh[k * 3 + c] * dL_dRGB[c];
            dL_ddir[0] += bx[k] * s;
            dL_ddir[1] += by[k] * s;
            dL_ddir[2] += bz[k] * s;
        }
}

void oracle_preprocess_backward(const OracleCam* cam, const float* means, const float* cov3d, const float* sh,
                                const int* radii, const uint8_t* clamped, const float* dL_dmean2D,
                                const float* dL_dconic, const float* dL_dcolor, const float* dL_ddepth,
                                float* dL_dmeans, float* dL_dcov3d, float* dL_dsh, double* dL_dcam) {
    /* dL_dcam (optional, 35 doubles): gradient w.r.t. the camera tensors of the settings, laid out
     * viewmatrix [4,4] | projmatrix [4,4] | campos [3] -- the opt-in extension for BASELINE config 3
     * ("through the projection into camera pose").  With row-vector maths t = [p,1].V[:, :3],
     * hom = [p,1].M and Tm = J.Rw (Rw[m][k] = V[k][m]):
     *   dL/dV[i][j] += p_i dL/dt_j  (i = 0..3, p_3 = 1)     dL/dV[k][m] += sum_r dL/dTm[r][k] J[r][m]
     *   dL/dM[i][j] += p_i dL/dhom_j (j = 0, 1, 3)          dL/dcampos  -= dL/d(p - campos)
     * Accumulated in double, per thread, merged at the end (deterministic up to the merge order). */
    const int P = cam->P;
    const int K = (cam->deg + 1) * (cam->deg + 1);
    const float* V = cam->view;
    const float* M = cam->proj;
    if (dL_dcam) memset(dL_dcam, 0, 35 * sizeof(double));
#pragma omp parallel
    {
    double camacc[35];
    memset(camacc, 0, sizeof(camacc));
#pragma omp for schedule(static)
    for (int i = 0; i < P; ++i) {
        for (int k = 0; k < 3; ++k) dL_dmeans[3 * i + k] = 0.0f;
        for (int k = 0; k < 6; ++k) dL_dcov3d[6 * i + k] = 0.0f;
        if (dL_dsh) memset(dL_dsh + (size_t)i * K * 3, 0, (size_t)K * 3 * sizeof(float));
        if (radii[i] <= 0) continue;
        Geo g;
        if (!geometry(cam, means + 3 * i, cov3d + 6 * i, &g)) continue;
        float dmean[3] = {0, 0, 0};

        /* conic -> cov2D */
        float gA = dL_dconic[3 * i], gB = dL_dconic[3 * i + 1], gC = dL_dconic[3 * i + 2];
        float a = g.a, b = g.b, c = g.c, det = g.det;
        float k2 = 1.0f / (det * det + 0.0000001f);
        float dL_da = k2 * (-c * c * gA + 2.0f * b * c * gB + (det - a * c) * gC);
        float dL_dc = k2 * (-a * a * gC + 2.0f * a * b * gB + (det - a * c) * gA);
        float dL_db = k2 * 2.0f * (b * c * gA - (det + 2.0f * b * b) * gB + a * b * gC);
        const float(*Tm)[3] = g.Tm;
        /* cov2D -> cov3D (xx,xy,xz,yy,yz,zz) */
        float* dS = dL_dcov3d + 6 * i;
        dS[0] = Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc;
        dS[3] = Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc;
        dS[5] = Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc;
        dS[1] = 2.0f * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db +
                2.0f * Tm[1][0] * Tm[1][1] * dL_dc;
        dS[2] = 2.0f * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db +
                2.0f * Tm[1][0] * Tm[1][2] * dL_dc;
        dS[4] = 2.0f * Tm[0][1] * Tm[0][2] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db +
                2.0f * Tm[1][1] * Tm[1][2] * dL_dc;
        /* cov2D -> Tm -> J -> t -> mean */
        const float* cv = cov3d + 6 * i;
        float S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
        float TS[2][3];
        for (int r = 0; r < 2; ++r)
            for (int k = 0; k < 3; ++k) TS[r][k] = Tm[r][0] * S[0][k] + Tm[r][1] * S[1][k] + Tm[r][2] * S[2][k];
        float dTm[2][3];
        for (int k = 0; k < 3; ++k) {
            dTm[0][k] = 2.0f * dL_da * TS[0][k] + dL_db * TS[1][k];
            dTm[1][k] = dL_db * TS[0][k] + 2.0f * dL_dc * TS[1][k];
        }
        /* J = Tm-contraction with world->camera rotation rows: dL_dJ[r][m] = sum_k dTm[r][k] * Rw[m][k], Rw[m][k]=V[4k+m] */
        float dJ00 = dTm[0][0] * V[0] + dTm[0][1] * V[4] + dTm[0][2] * V[8];
        float dJ02 = dTm[0][0] * V[2] + dTm[0][1] * V[6] + dTm[0][2] * V[10];
        float dJ11 = dTm[1][0] * V[1] + dTm[1][1] * V[5] + dTm[1][2] * V[9];
        float dJ12 = dTm[1][0] * V[2] + dTm[1][1] * V[6] + dTm[1][2] * V[10];
        float limx = 1.3f * cam->tanfovx, limy = 1.3f * cam->tanfovy;
        float xmul = (g.txtz < -limx || g.txtz > limx) ? 0.0f : 1.0f;
        float ymul = (g.tytz < -limy || g.tytz > limy) ? 0.0f : 1.0f;
        float z1 = 1.0f / g.tz, z2 = z1 * z1, z3 = z2 * z1;
        float dtx = xmul * -g.fx * z2 * dJ02;
        float dty = ymul * -g.fy * z2 * dJ12;
        float dtz = -g.fx * z2 * dJ00 - g.fy * z2 * dJ11 + (2.0f * g.fx * g.cx) * z3 * dJ02 +
                    (2.0f * g.fy * g.cy) * z3 * dJ12;
        if (dL_ddepth) dtz += dL_ddepth[i];
        for (int k = 0; k < 3; ++k) dmean[k] += V[4 * k + 0] * dtx + V[4 * k + 1] * dty + V[4 * k + 2] * dtz;

        /* NDC mean -> mean3D through the full projection */
        float mw = g.pw;
        float mul1 = g.hx * mw * mw, mul2 = g.hy * mw * mw;
        float g2x = dL_dmean2D[2 * i], g2y = dL_dmean2D[2 * i + 1];
        for (int k = 0; k < 3; ++k)
            dmean[k] += (M[4 * k + 0] * mw - M[4 * k + 3] * mul1) * g2x + (M[4 * k + 1] * mw - M[4 * k + 3] * mul2) * g2y;
        if (dL_dcam) {
            const double ph[4] = {means[3 * i], means[3 * i + 1], means[3 * i + 2], 1.0};
            const double dt[3] = {dtx, dty, dtz};
            const double J00 = g.fx * z1, J02 = -g.fx * g.cx * z2, J11 = g.fy * z1, J12 = -g.fy * g.cy * z2;
            const double dhx = (double)g2x * mw, dhy = (double)g2y * mw;
            const double dhw = -((double)g.hx * g2x + (double)g.hy * g2y) * mw * mw;
            for (int r = 0; r < 4; ++r) {
                for (int j = 0; j < 3; ++j) camacc[4 * r + j] += ph[r] * dt[j];
                camacc[16 + 4 * r + 0] += ph[r] * dhx;
                camacc[16 + 4 * r + 1] += ph[r] * dhy;
                camacc[16 + 4 * r + 3] += ph[r] * dhw;
            }
            for (int k = 0; k < 3; ++k) {
                camacc[4 * k + 0] += (double)dTm[0][k] * J00;
                camacc[4 * k + 1] += (double)dTm[1][k] * J11;
                camacc[4 * k + 2] += (double)dTm[0][k] * J02 + (double)dTm[1][k] * J12;
            }
        }

        /* colour -> SH and view direction */
        if (dL_dsh) {
            float dx = means[3 * i] - cam->campos[0], dy = means[3 * i + 1] - cam->campos[1],
                  dz = means[3 * i + 2] - cam->campos[2];
            float len2 = dx * dx + dy * dy + dz * dz;
            float inv = 1.0f / sqrtf(len2);
            float dRGB[3], ddir[3];
            for (int c2 = 0; c2 < 3; ++c2) dRGB[c2] = clamped[3 * i + c2] ? 0.0f : dL_dcolor[3 * i + c2];
            sh_backward(cam->deg, sh + (size_t)i * K * 3, dx * inv, dy * inv, dz * inv, dRGB,
                        dL_dsh + (size_t)i * K * 3, ddir);
            /* through normalize(): (I |v|^2 - v v^T) / |v|^3 . ddir */
            float inv3 = inv * inv * inv;
            float dot = dx * ddir[0] + dy * ddir[1] + dz * ddir[2];
            const float m0 = (len2 * ddir[0] - dx * dot) * inv3, m1 = (len2 * ddir[1] - dy * dot) * inv3,
                        m2 = (len2 * ddir[2] - dz * dot) * inv3;
            dmean[0] += m0;
            dmean[1] += m1;
            dmean[2] += m2;
            if (dL_dcam) camacc[32] -= m0, camacc[33] -= m1, camacc[34] -= m2;
        }
        for (int k = 0; k < 3; ++k) dL_dmeans[3 * i + k] = dmean[k];
    }
    if (dL_dcam) {
#pragma omp critical
        for (int k = 0; k < 35; ++k) dL_dcam[k] += camacc[k];
    }
    }  /* omp parallel */
}

/* A.? mark_visible: the frustum test alone (upstream checkFrustum; unused by GGRt). */
void oracle_mark_visible(const OracleCam* cam, const float* means, uint8_t* visible) {
    const float* V = cam->view;
    for (int i = 0; i < cam->P; ++i) {
        const float* p = means + 3 * i;
        float tz = ((V[2] * p[0] + V[6] * p[1]) + V[10] * p[2]) + V[14];
        visible[i] = (uint8_t)(tz > NEAR_CULL);
    }
}
