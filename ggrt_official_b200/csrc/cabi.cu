// extern "C" entry points of libggrt_raster.so (include/ggrt_raster.h) and buffer layout.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h>  // header-only: resolves the tools library lazily, a no-op without a profiler attached

#include "common.cuh"

namespace ggrt {

static thread_local char g_err[512] = "";

struct NvtxRange {  // RAII NVTX range around a launch
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static const char* const kStageNames[GGRT_STAGE_COUNT] = {"geometry",       "scan_tiles",      "color",
                                                          "emit",           "sort_tiles",      "render_forward",
                                                          "render_backward", "preprocess_backward"};

bool pdl_enabled(cudaStream_t s) {
    static const bool on = [] {
        const char* e = getenv("GGRT_RASTER_PDL");
        return !(e && e[0] == '0');
    }();
    if (!on) return false;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st == cudaStreamCaptureStatusActive;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what, int debug, cudaStream_t s) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return GGRT_ERR_CUDA;
    }
    return GGRT_OK;
}

// ---- optional per-kernel timing -------------------------------------------------------
struct Profiler {
    bool on = false;
    bool made = false;
    bool used[GGRT_STAGE_COUNT] = {};
    cudaEvent_t a[GGRT_STAGE_COUNT], b[GGRT_STAGE_COUNT];
};
static thread_local Profiler g_prof;

struct StageTimer {  // RAII: records start/stop events around one launch when profiling is on
    int stage;
    cudaStream_t s;
    StageTimer(int stage_, cudaStream_t s_) : stage(stage_), s(s_) {
        nvtxRangePushA(kStageNames[stage]);  // one NVTX range per kernel launch (nsys / ncu --nvtx), SURVEY.md section 5
        if (!g_prof.on) return;
        if (!g_prof.made) {
            for (int i = 0; i < GGRT_STAGE_COUNT; ++i) cudaEventCreate(&g_prof.a[i]), cudaEventCreate(&g_prof.b[i]);
            g_prof.made = true;
        }
        cudaEventRecord(g_prof.a[stage], s);
    }
    ~StageTimer() {
        nvtxRangePop();
        if (!g_prof.on) return;
        cudaEventRecord(g_prof.b[stage], s);
        g_prof.used[stage] = true;
    }
};

// ---- colour evaluation overlapped with binning ------------------------------------------
// The SH colour kernel is an HBM stream (~25 us at C2) that only the render kernel consumes, while the
// tile scan / emit / sort kernels between them are latency- and L2-bound and leave HBM idle.  `prepare`
// therefore forks the colour kernel onto an internal per-(thread, device) side stream right after the
// geometry kernel and `render` joins it in front of the render kernel.  The only state kept is this stream
// and its two events; GGRT_RASTER_OVERLAP=0 (or profiling / debug mode) keeps everything on the caller's stream.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t fork_b = nullptr, join_b = nullptr;  // backward: the dL/dsh writer beside the per-Gaussian kernel
    bool pending = false;
    unsigned long long capture_id = 0;  // stream-capture sequence the join event was last recorded in (0: none)
};

// id of the CUDA-graph capture `s` is part of (0 when it is not capturing).  An event recorded inside a capture can
// only be waited on inside the same capture and vice versa, so the fork/join of the colour kernel is matched by id.
static unsigned long long capture_id_of(cudaStream_t s) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo(s, &st, &id) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return st == cudaStreamCaptureStatusActive ? (id ? id : 1ull) : 0ull;
}
constexpr int MAX_DEVICES = 64;
static thread_local SideStream g_side[MAX_DEVICES];

static bool overlap_enabled() {
    static const bool on = [] {
        const char* e = getenv("GGRT_RASTER_OVERLAP");
        return !(e && e[0] == '0');
    }();
    return on;
}

// returns the side stream of the current device, or nullptr when it cannot be used
static SideStream* side_stream() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return nullptr;
    SideStream* ss = &g_side[dev];
    if (!ss->stream) {
        if (cudaStreamCreateWithFlags(&ss->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ss->fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ss->join, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ss->fork_b, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ss->join_b, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            ss->stream = nullptr;
            return nullptr;
        }
    }
    return ss;
}

void compute_layout(int P, int H, int W, long long N, GgrtRasterLayout* L) {
    memset(L, 0, sizeof(*L));
    const size_t p = (size_t)(P > 0 ? P : 0);
    size_t off = 0;
    L->geom_rec0 = off; off = align_up(off + p * sizeof(float4));
    L->geom_rec1 = off; off = align_up(off + p * sizeof(float4));
    L->geom_rec2 = off; off = align_up(off + p * sizeof(float4));
    L->geom_rect = off; off = align_up(off + p * sizeof(ushort4));
    L->geom_tiles = off; off = align_up(off + p * sizeof(uint32_t));
    L->geom_flags = off; off = align_up(off + p * sizeof(uint8_t));
    L->geom_ranks = off; off = align_up(off + p * sizeof(uint4));
    L->geom_jac = off; off = align_up(off + 9 * p * sizeof(float));
    L->geom_bytes = off > 0 ? off : 256;

    const size_t gx = (size_t)(W + TILE - 1) / TILE, gy = (size_t)(H + TILE - 1) / TILE, T = gx * gy;
    const size_t px = (size_t)H * W;
    off = 0;
    L->img_counts = off; off = off + T * SUBS * sizeof(uint32_t);
    L->img_partials = off; off = align_up(off + 2 * ((T + SCAN_BLOCK - 1) / SCAN_BLOCK) * sizeof(unsigned long long));
    L->img_cursor = off; off = align_up(off + T * SUBS * sizeof(uint32_t));
    L->img_starts = off; off = align_up(off + (T + 1) * sizeof(uint32_t));
    L->img_header = off; off = align_up(off + 4 * sizeof(uint32_t));
    L->img_final_T = off; off = align_up(off + px * sizeof(float));
    L->img_ncontrib = off; off = align_up(off + px * sizeof(uint32_t));
    L->img_bytes = off;

    const size_t n = (size_t)(N > 0 ? N : 0);
    off = 0;
    L->bin_keys = off; off = align_up(off + n * sizeof(unsigned long long));
    L->bin_points = off; off = align_up(off + n * sizeof(uint32_t));
    L->bin_masks = off; off = align_up(off + n * sizeof(uint8_t));
    L->bin_bytes = off > 0 ? off : 256;
}

GeomPtrs geom_ptrs(void* base, int P) {
    GgrtRasterLayout L;
    compute_layout(P, 0, 0, 0, &L);
    char* b = static_cast<char*>(base);
    GeomPtrs g;
    g.rec0 = reinterpret_cast<float4*>(b + L.geom_rec0);
    g.rec1 = reinterpret_cast<float4*>(b + L.geom_rec1);
    g.rec2 = reinterpret_cast<float4*>(b + L.geom_rec2);
    g.rect = reinterpret_cast<ushort4*>(b + L.geom_rect);
    g.tiles = reinterpret_cast<uint32_t*>(b + L.geom_tiles);
    g.flags = reinterpret_cast<uint8_t*>(b + L.geom_flags);
    g.ranks = reinterpret_cast<uint4*>(b + L.geom_ranks);
    g.jac = reinterpret_cast<float*>(b + L.geom_jac);
    g.jac_plane = (size_t)(P > 0 ? P : 0);
    return g;
}

ImagePtrs image_ptrs(void* base, int H, int W) {
    GgrtRasterLayout L;
    compute_layout(0, H, W, 0, &L);
    char* b = static_cast<char*>(base);
    ImagePtrs im;
    im.counts = reinterpret_cast<uint32_t*>(b + L.img_counts);
    im.partials = reinterpret_cast<unsigned long long*>(b + L.img_partials);
    im.starts = reinterpret_cast<uint32_t*>(b + L.img_starts);
    im.cursor = reinterpret_cast<uint32_t*>(b + L.img_cursor);
    im.header = reinterpret_cast<uint32_t*>(b + L.img_header);
    im.final_T = reinterpret_cast<float*>(b + L.img_final_T);
    im.n_contrib = reinterpret_cast<uint32_t*>(b + L.img_ncontrib);
    return im;
}

BinPtrs bin_ptrs(void* base, long long N) {
    GgrtRasterLayout L;
    compute_layout(0, 0, 0, N, &L);
    char* b = static_cast<char*>(base);
    BinPtrs p;
    p.keys = reinterpret_cast<unsigned long long*>(b + L.bin_keys);
    p.points = reinterpret_cast<uint32_t*>(b + L.bin_points);
    p.masks = reinterpret_cast<uint8_t*>(b + L.bin_masks);
    return p;
}

static int make_view(const GgrtRasterSettings* s, const GgrtRasterInputLayout* lay, int P, View* v) {
    if (s == nullptr) {
        set_error("settings is NULL");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (P < 0 || s->image_height <= 0 || s->image_width <= 0) {
        set_error("bad sizes: P=%d H=%d W=%d", P, s->image_height, s->image_width);
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (s->sh_degree < 0 || s->sh_degree > 4) {
        set_error("sh_degree %d outside 0..4", s->sh_degree);
        return GGRT_ERR_UNSUPPORTED;
    }
    if (s->device_params == nullptr && (!(s->tanfovx > 0.f) || !(s->tanfovy > 0.f))) {
        set_error("tanfov must be positive");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (s->aux_mode != 0 && s->aux_mode != 1) {
        set_error("aux_mode %d unknown", s->aux_mode);
        return GGRT_ERR_UNSUPPORTED;
    }
    if (!s->viewmatrix || !s->projmatrix || !s->campos || !s->bg) {
        set_error("viewmatrix / projmatrix / campos / bg must be device pointers");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    v->W = s->image_width;
    v->H = s->image_height;
    v->gx = (v->W + TILE - 1) / TILE;
    v->gy = (v->H + TILE - 1) / TILE;
    if (v->gx > 65535 || v->gy > 65535) {
        set_error("image too large");
        return GGRT_ERR_UNSUPPORTED;
    }
    v->P = P;
    v->deg = s->sh_degree;
    v->K = (s->sh_degree + 1) * (s->sh_degree + 1);
    v->tanfovx = s->tanfovx;
    v->tanfovy = s->tanfovy;
    v->fx = (float)v->W / (2.0f * s->tanfovx);  // same float expression as the oracle
    v->fy = (float)v->H / (2.0f * s->tanfovy);
    v->scale = 1.0f;
    v->cov_stride = 6;
    v->sh_ks = 3, v->sh_cs = 1;
    if (lay != nullptr) {
        if (!(lay->scene_scale > 0.f)) {
            set_error("scene_scale must be positive");
            return GGRT_ERR_INVALID_ARGUMENT;
        }
        v->scale = lay->scene_scale;
        v->cov_stride = lay->cov_full3x3 ? 9 : 6;
        if (lay->sh_channel_major) v->sh_ks = 1, v->sh_cs = v->K;
    }
    v->view = s->viewmatrix;
    v->proj = s->projmatrix;
    v->campos = s->campos;
    v->bg = s->bg;
    v->dparams = s->device_params;
    v->aux_mode = s->aux_mode;
    if (s->device_params != nullptr) {  // the kernels read tanfov / scene scale themselves (resolve_device_params)
        v->tanfovx = v->tanfovy = v->fx = v->fy = 0.0f;
    }
    return GGRT_OK;
}

}  // namespace ggrt

using namespace ggrt;

#define GGRT_TRY(expr)                    \
    do {                                  \
        const int rc_ = (expr);           \
        if (rc_ != GGRT_OK) return rc_;   \
    } while (0)

extern "C" {

int ggrt_raster_abi_version(void) { return GGRT_RASTER_ABI_VERSION; }

const char* ggrt_raster_last_error(void) { return g_err; }

int ggrt_raster_layout(int32_t P, int32_t H, int32_t W, int64_t N, GgrtRasterLayout* out) {
    if (!out || P < 0 || H < 0 || W < 0 || N < 0) {
        set_error("ggrt_raster_layout: bad argument");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    compute_layout(P, H, W, N, out);
    return GGRT_OK;
}

size_t ggrt_raster_geom_bytes(int32_t P) {
    GgrtRasterLayout L;
    compute_layout(P, 0, 0, 0, &L);
    return L.geom_bytes;
}

size_t ggrt_raster_image_bytes(int32_t H, int32_t W) {
    GgrtRasterLayout L;
    compute_layout(0, H, W, 0, &L);
    return L.img_bytes;
}

size_t ggrt_raster_binning_bytes(int64_t N) {
    GgrtRasterLayout L;
    compute_layout(0, 0, 0, N, &L);
    return L.bin_bytes;
}

int ggrt_raster_forward_prepare(const GgrtRasterSettings* settings, const GgrtRasterInputLayout* layout, int32_t P,
                                const float* means3D,
                                const float* cov3D_precomp, const float* opacities, const float* shs,
                                const float* colors_precomp, const float* aux, int32_t* radii, void* geom_buffer,
                                void* image_buffer, uint32_t* counts_host, ggrt_stream_t stream) {
    View v;
    GGRT_TRY(make_view(settings, layout, P, &v));
    if (P > 0 && (shs == nullptr) == (colors_precomp == nullptr)) {
        set_error("exactly one of shs / colors_precomp must be given");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (!geom_buffer || !image_buffer || (P > 0 && (!means3D || !cov3D_precomp || !opacities || !radii))) {
        set_error("forward_prepare: NULL buffer");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int dbg = settings->debug;
    GeomPtrs g = geom_ptrs(geom_buffer, P);
    ImagePtrs im = image_ptrs(image_buffer, v.H, v.W);
    if (cudaMemsetAsync(im.counts, 0, reinterpret_cast<char*>(im.cursor) - reinterpret_cast<char*>(im.counts), s) != cudaSuccess)
        return check_launch("memset tile counts", 0, s);
    for (int i = 0; i < GGRT_STAGE_COUNT; ++i) g_prof.used[i] = false;
    { StageTimer t_(GGRT_STAGE_GEOMETRY, s); launch_geometry(v, means3D, cov3D_precomp, opacities, radii, g, im, s); }
    GGRT_TRY(check_launch("geometry", dbg, s));
    // colour evaluation needs the geometry kernel's radii / depths but nothing of the binning: fork it onto the
    // side stream so it streams the SH table from HBM while scan / emit / sort run (joined in forward_render).
    // Measured at C2: 0.387 -> 0.377 ms per step.  (Forking it before the geometry kernel, with the colour kernel
    // evaluating every Gaussian and deriving the depth itself, was measured too and is no faster: 0.378 ms.)
    SideStream* ss = (overlap_enabled() && !g_prof.on && !dbg && P > 0) ? side_stream() : nullptr;
    if (ss && cudaEventRecord(ss->fork, s) == cudaSuccess && cudaStreamWaitEvent(ss->stream, ss->fork, 0) == cudaSuccess) {
        nvtxRangePushA("color (side stream)");
        launch_color(v, means3D, shs, colors_precomp, aux, radii, g, ss->stream);
        nvtxRangePop();
        GGRT_TRY(check_launch("color", 0, ss->stream));
        // the backward's scratch is zeroed here, off the critical path (GgrtRasterSettings.zero_scratch)
        if (settings->zero_scratch &&
            cudaMemsetAsync(settings->zero_scratch, 0, (size_t)P * GRAD_STRIDE * sizeof(float), ss->stream) != cudaSuccess)
            return check_launch("memset grad scratch", 0, ss->stream);
        if (cudaEventRecord(ss->join, ss->stream) != cudaSuccess) return check_launch("color join", 0, s);
        ss->pending = true;
        ss->capture_id = capture_id_of(s);
    } else {
        ss = nullptr;
    }
    // {N, max pairs per tile} goes straight into the caller's mapped pinned host memory from the scan kernel
    { StageTimer t_(GGRT_STAGE_SCAN_TILES, s); launch_scan_tiles(v, im, counts_host, s); }
    GGRT_TRY(check_launch("scan_tiles", dbg, s));
    if (!ss) {
        { StageTimer t_(GGRT_STAGE_COLOR, s); launch_color(v, means3D, shs, colors_precomp, aux, radii, g, s); }
        GGRT_TRY(check_launch("color", dbg, s));
        if (settings->zero_scratch && P > 0 &&
            cudaMemsetAsync(settings->zero_scratch, 0, (size_t)P * GRAD_STRIDE * sizeof(float), s) != cudaSuccess)
            return check_launch("memset grad scratch", 0, s);
    }
    return GGRT_OK;
}

int ggrt_raster_forward_render(const GgrtRasterSettings* settings, int32_t P, int64_t num_rendered,
                               uint32_t max_tile_pairs, int32_t rescan, const void* geom_buffer, void* binning_buffer,
                               void* image_buffer, float* out_color, float* out_depth, ggrt_stream_t stream) {
    View v;
    GGRT_TRY(make_view(settings, nullptr, P, &v));
    if (!geom_buffer || !image_buffer || !out_color || !out_depth || num_rendered < 0 ||
        (num_rendered > 0 && !binning_buffer)) {
        set_error("forward_render: bad argument");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (num_rendered > 0xffffffffLL) {
        set_error("more than 2^32 tile-Gaussian pairs");
        return GGRT_ERR_UNSUPPORTED;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int dbg = settings->debug;
    GeomPtrs g = geom_ptrs(const_cast<void*>(geom_buffer), P);
    ImagePtrs im = image_ptrs(image_buffer, v.H, v.W);
    BinPtrs b = bin_ptrs(binning_buffer, num_rendered);
    const uint32_t capacity = (uint32_t)num_rendered;
    if (rescan) {  // a previous speculative attempt consumed the emit cursors
        { StageTimer t_(GGRT_STAGE_SCAN_TILES, s); launch_scan_tiles(v, im, nullptr, s); }
        GGRT_TRY(check_launch("scan_tiles", dbg, s));
    }
    if (num_rendered > 0) {
        { StageTimer t_(GGRT_STAGE_EMIT, s); launch_emit(v, g, im, b, capacity, s); }
        GGRT_TRY(check_launch("emit", dbg, s));
        { StageTimer t_(GGRT_STAGE_SORT_TILES, s); launch_sort_tiles(v, im, b, max_tile_pairs, capacity, s); }
        GGRT_TRY(check_launch("sort_tiles", dbg, s));
    }
    {  // join the colour kernel that `prepare` forked (a later event of the in-order side stream covers earlier ones)
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < MAX_DEVICES && g_side[dev].pending &&
            g_side[dev].capture_id == capture_id_of(s)) {
            // `pending` is never cleared: renders of interleaved prepares on other streams must wait as well, and
            // waiting on an event that has already completed costs nothing on the device.  (An event last recorded
            // inside another graph capture -- or outside, while `s` is capturing -- belongs to a frame that is not
            // this one: nothing to wait for, and waiting across a capture boundary is an error.)
            if (cudaStreamWaitEvent(s, g_side[dev].join, 0) != cudaSuccess) return check_launch("color join", 0, s);
        }
    }
    { StageTimer t_(GGRT_STAGE_RENDER_FORWARD, s); launch_render_forward(v, g, im, b, capacity, out_color, out_depth, s); }
    GGRT_TRY(check_launch("render_forward", dbg, s));
    return GGRT_OK;
}

int ggrt_raster_join(ggrt_stream_t stream) {
    // Makes `stream` wait for the colour kernel a forward_prepare of this thread forked onto the internal side
    // stream.  forward_render does this itself; a caller that abandons a frame between prepare and render (an
    // allocation failure, say) calls it so that the buffers prepare was given can be released safely.
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return check_launch("join", 0, nullptr);
    if (g_side[dev].pending && g_side[dev].capture_id == capture_id_of(static_cast<cudaStream_t>(stream)) &&
        cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), g_side[dev].join, 0) != cudaSuccess)
        return check_launch("join", 0, static_cast<cudaStream_t>(stream));
    return GGRT_OK;
}

int ggrt_raster_backward(const GgrtRasterSettings* settings, const GgrtRasterInputLayout* layout, int32_t P,
                         int64_t num_rendered, const float* means3D,
                         const float* cov3D_precomp, const float* shs, const int32_t* radii, const void* geom_buffer,
                         const void* binning_buffer, const void* image_buffer, const float* dL_dout_color,
                         const float* dL_dout_aux, float* grad_scratch, float* dL_dmeans2D, float* dL_dopacity,
                         float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dcolors, float* dL_daux,
                         float* dL_dcamera, const GgrtRasterGradSinks* color_sinks, ggrt_stream_t stream) {
    View v;
    GGRT_TRY(make_view(settings, layout, P, &v));
    if (P == 0) return GGRT_OK;
    if (!means3D || !cov3D_precomp || !radii || !geom_buffer || !image_buffer || !dL_dout_color || !grad_scratch ||
        !dL_dmeans2D || !dL_dopacity || !dL_dmeans3D || !dL_dcov3D || (num_rendered > 0 && !binning_buffer)) {
        set_error("backward: NULL buffer");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (settings->aux_mode == 1 ? dL_daux != nullptr : (dL_dout_aux != nullptr) != (dL_daux != nullptr)) {
        set_error("backward: dL_dout_aux and dL_daux go together (aux_mode 1: dL_daux must be NULL, the gradient of the "
                  "depth channel flows into dL_dmeans3D)");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (dL_dout_aux == nullptr) v.aux_mode = 0;  // no gradient for the channel: the per-Gaussian kernel need not know it
    // with shs: dL_dsh (full SH gradient) or dL_dcolors / color_sinks (compact mode, see the header)
    const bool have_sinks = color_sinks != nullptr;
    if ((dL_dsh != nullptr) == (dL_dcolors != nullptr || have_sinks) || (shs == nullptr && (dL_dsh != nullptr || have_sinks)) ||
        (have_sinks && dL_dcolors != nullptr)) {
        set_error("backward: pass exactly one of dL_dsh (needs shs) / dL_dcolors / color_sinks (needs shs)");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    ColorSinks sinks;
    memset(&sinks, 0, sizeof(sinks));
    if (have_sinks) {
        if (color_sinks->count < 1 || color_sinks->count > GGRT_RASTER_MAX_MERGE_VIEWS ||
            (color_sinks->multimem && color_sinks->count != 1)) {
            set_error("backward: color_sinks->count must be 1..%d (exactly 1 with multimem)", GGRT_RASTER_MAX_MERGE_VIEWS);
            return GGRT_ERR_INVALID_ARGUMENT;
        }
        for (int k = 0; k < color_sinks->count; ++k) {
            if (!color_sinks->ptr[k] || (reinterpret_cast<uintptr_t>(color_sinks->ptr[k]) & 15)) {
                set_error("backward: color_sinks->ptr[%d] must be a 16-byte aligned device pointer", k);
                return GGRT_ERR_INVALID_ARGUMENT;
            }
            sinks.ptr[k] = color_sinks->ptr[k];
        }
        sinks.n = color_sinks->count, sinks.multimem = color_sinks->multimem != 0, sinks.with_campos = 1;
        if (color_sinks->epoch != nullptr) {  // in-kernel step signalling
            const int na = color_sinks->arrive_count;
            if (!color_sinks->done_counter || na < 1 || na > GGRT_RASTER_MAX_MERGE_VIEWS ||
                (color_sinks->multimem && na != 1) || color_sinks->parity_stride < 0 || (color_sinks->parity_stride & 3)) {
                set_error("backward: signalling needs done_counter, 1..%d arrival counters (1 with multimem) and a "
                          "parity_stride that is a non-negative multiple of 4 floats", GGRT_RASTER_MAX_MERGE_VIEWS);
                return GGRT_ERR_INVALID_ARGUMENT;
            }
            for (int k = 0; k < na; ++k) {
                if (!color_sinks->arrive[k] || (reinterpret_cast<uintptr_t>(color_sinks->arrive[k]) & 3)) {
                    set_error("backward: color_sinks->arrive[%d] must be a 4-byte aligned device pointer", k);
                    return GGRT_ERR_INVALID_ARGUMENT;
                }
                sinks.arrive[k] = color_sinks->arrive[k];
            }
            sinks.epoch = color_sinks->epoch, sinks.done = color_sinks->done_counter;
            sinks.parity_stride = color_sinks->parity_stride, sinks.n_arrive = na;
        }
    } else if (shs != nullptr && dL_dcolors != nullptr) {
        sinks.ptr[0] = dL_dcolors, sinks.n = 1;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int dbg = settings->debug;
    GeomPtrs g = geom_ptrs(const_cast<void*>(geom_buffer), P);
    ImagePtrs im = image_ptrs(const_cast<void*>(image_buffer), v.H, v.W);
    BinPtrs b = bin_ptrs(const_cast<void*>(binning_buffer), num_rendered);
    if (settings->zero_scratch != grad_scratch &&  // (else forward_prepare has zeroed it already)
        cudaMemsetAsync(grad_scratch, 0, (size_t)P * GRAD_STRIDE * sizeof(float), s) != cudaSuccess)
        return check_launch("memset grad scratch", 0, s);
    if (dL_dcamera && cudaMemsetAsync(dL_dcamera, 0, 35 * sizeof(float), s) != cudaSuccess)
        return check_launch("memset camera gradient", 0, s);
    if (num_rendered > 0) {
        { StageTimer t_(GGRT_STAGE_RENDER_BACKWARD, s); launch_render_backward(v, g, im, b, dL_dout_color, dL_dout_aux, grad_scratch, s); }
        GGRT_TRY(check_launch("render_backward", dbg, s));
    }
    {
        // with a full SH gradient to write, its streaming kernel runs on the side stream beside the per-Gaussian kernel
        SideStream* ss = (overlap_enabled() && !g_prof.on && !dbg && shs != nullptr && dL_dsh != nullptr) ? side_stream() : nullptr;
        if (ss && !(cudaEventRecord(ss->fork_b, s) == cudaSuccess && cudaStreamWaitEvent(ss->stream, ss->fork_b, 0) == cudaSuccess)) {
            cudaGetLastError();
            ss = nullptr;
        }
        StageTimer t_(GGRT_STAGE_PREPROCESS_BACKWARD, s);
        launch_preprocess_backward(v, means3D, cov3D_precomp, shs, radii, g, grad_scratch, dL_dmeans2D, dL_dopacity,
                                   dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dcolors, dL_daux, dL_dcamera, sinks, s,
                                   ss ? ss->stream : nullptr);
        if (ss && !(cudaEventRecord(ss->join_b, ss->stream) == cudaSuccess && cudaStreamWaitEvent(s, ss->join_b, 0) == cudaSuccess))
            return check_launch("dL/dsh join", 0, s);
    }
    GGRT_TRY(check_launch("preprocess_backward", dbg, s));
    return GGRT_OK;
}

int ggrt_raster_sh_gradient_merge(int32_t P, int32_t sh_degree, const GgrtRasterInputLayout* layout,
                                  const float* means3D, int32_t num_views, const float* const* drgb_views_host,
                                  const float* const* campos_views_host, float* dL_dsh, ggrt_stream_t stream) {
    if (P < 0 || sh_degree < 0 || sh_degree > 4 || num_views < 1 || num_views > GGRT_RASTER_MAX_MERGE_VIEWS) {
        set_error("sh_gradient_merge: bad sizes (P=%d, sh_degree=%d, num_views=%d, at most %d views)", P, sh_degree,
                  num_views, GGRT_RASTER_MAX_MERGE_VIEWS);
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (!drgb_views_host || !campos_views_host || (P > 0 && (!means3D || !dL_dsh))) {
        set_error("sh_gradient_merge: NULL buffer");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    for (int v = 0; v < num_views; ++v)
        if (!drgb_views_host[v] || !campos_views_host[v]) {
            set_error("sh_gradient_merge: NULL pointer for view %d", v);
            return GGRT_ERR_INVALID_ARGUMENT;
        }
    float scale = 1.0f;
    bool cmajor = false;
    if (layout) {
        if (!(layout->scene_scale > 0.f)) {
            set_error("scene_scale must be positive");
            return GGRT_ERR_INVALID_ARGUMENT;
        }
        scale = layout->scene_scale;
        cmajor = layout->sh_channel_major != 0;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NvtxRange nvtx_("sh_gradient_merge");
    launch_sh_gradient_merge(P, sh_degree, scale, cmajor, means3D, num_views, drgb_views_host, campos_views_host,
                             dL_dsh, MergeSignal(), s);
    return check_launch("sh_gradient_merge", 0, s);
}

int ggrt_raster_sh_gradient_merge_signalled(int32_t P, int32_t sh_degree, const GgrtRasterInputLayout* layout,
                                            const float* means3D, int32_t world, const float* slots,
                                            int64_t slot_stride, int64_t parity_stride, const uint32_t* epoch,
                                            const uint32_t* arrive, float* dL_dsh, ggrt_stream_t stream) {
    if (P < 0 || sh_degree < 0 || sh_degree > 4 || world < 1 || world > GGRT_RASTER_MAX_MERGE_VIEWS ||
        slot_stride < 3LL * (P + 1) || parity_stride < slot_stride * world) {
        set_error("sh_gradient_merge_signalled: bad sizes (P=%d, sh_degree=%d, world=%d, at most %d ranks; slots are "
                  "[P+1,3] floats)", P, sh_degree, world, GGRT_RASTER_MAX_MERGE_VIEWS);
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (!slots || !epoch || !arrive || (P > 0 && (!means3D || !dL_dsh))) {
        set_error("sh_gradient_merge_signalled: NULL buffer");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    float scale = 1.0f;
    bool cmajor = false;
    if (layout) {
        if (!(layout->scene_scale > 0.f)) {
            set_error("scene_scale must be positive");
            return GGRT_ERR_INVALID_ARGUMENT;
        }
        scale = layout->scene_scale;
        cmajor = layout->sh_channel_major != 0;
    }
    const float* drgb[GGRT_RASTER_MAX_MERGE_VIEWS];
    const float* campos[GGRT_RASTER_MAX_MERGE_VIEWS];
    for (int v = 0; v < world; ++v) drgb[v] = slots + (size_t)v * slot_stride, campos[v] = drgb[v] + 3 * (size_t)P;
    MergeSignal sig;
    sig.epoch = epoch, sig.arrive = arrive, sig.world = world, sig.parity_stride = parity_stride;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (P == 0) {
        set_error("sh_gradient_merge_signalled: P = 0 is not supported (the kernel is the step's wait)");
        return GGRT_ERR_UNSUPPORTED;
    }
    NvtxRange nvtx_("sh_gradient_merge (signalled)");
    launch_sh_gradient_merge(P, sh_degree, scale, cmajor, means3D, world, drgb, campos, dL_dsh, sig, s);
    return check_launch("sh_gradient_merge_signalled", 0, s);
}

int ggrt_raster_nvls_allreduce_f32(void* multicast_ptr, int64_t count, int32_t rank, int32_t world,
                                   ggrt_stream_t stream) {
    if (!multicast_ptr || count < 0 || (count & 3) || world < 1 || rank < 0 || rank >= world ||
        (reinterpret_cast<uintptr_t>(multicast_ptr) & 15)) {
        set_error("nvls_allreduce: bad argument (count must be a multiple of 4, pointer 16-byte aligned)");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NvtxRange nvtx_("nvls_allreduce");
    launch_nvls_allreduce(static_cast<float*>(multicast_ptr), count, rank, world, nullptr, nullptr, nullptr, nullptr,
                          nullptr, s);
    return check_launch("nvls_allreduce", 0, s);
}

int ggrt_raster_nvls_allreduce_signalled(void* multicast_ptr, int64_t count, int32_t rank, int32_t world,
                                         const uint32_t* epoch, const uint32_t* arrive_in, void* arrive_out_multicast,
                                         const uint32_t* arrive_out, uint32_t* done_counter, ggrt_stream_t stream) {
    if (!multicast_ptr || count < 0 || (count & 3) || world < 1 || rank < 0 || rank >= world ||
        (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) || !epoch || !arrive_in || !arrive_out_multicast || !arrive_out ||
        !done_counter) {
        set_error("nvls_allreduce_signalled: bad argument (count must be a multiple of 4, pointers non-NULL, data "
                  "pointer 16-byte aligned)");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NvtxRange nvtx_("nvls_allreduce (signalled)");
    launch_nvls_allreduce(static_cast<float*>(multicast_ptr), count, rank, world, epoch, arrive_in,
                          static_cast<uint32_t*>(arrive_out_multicast), arrive_out, done_counter, s);
    return check_launch("nvls_allreduce_signalled", 0, s);
}

int ggrt_raster_nvls_barrier(void* multicast_counter, const void* local_counter, uint32_t target, ggrt_stream_t stream) {
    if (!multicast_counter || !local_counter || (reinterpret_cast<uintptr_t>(multicast_counter) & 3) ||
        (reinterpret_cast<uintptr_t>(local_counter) & 3)) {
        set_error("nvls_barrier: counters must be 4-byte aligned device pointers");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    launch_nvls_barrier(static_cast<unsigned int*>(multicast_counter), static_cast<const unsigned int*>(local_counter),
                        target, s);
    return check_launch("nvls_barrier", 0, s);
}

static int check_adapter_params(const GgrtAdapterParams* p) {
    if (!p) {
        set_error("adapter: params is NULL");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (p->num_views < 0 || p->rays_per_view < 0 || p->samples_per_ray < 0 || p->image_height <= 0 || p->image_width <= 0) {
        set_error("adapter: bad sizes (views=%d rays/view=%d samples/ray=%d image=%dx%d)", p->num_views, p->rays_per_view,
                  p->samples_per_ray, p->image_width, p->image_height);
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    if (p->sh_degree < 0 || p->sh_degree > 4) {
        set_error("adapter: sh_degree %d outside 0..4", p->sh_degree);
        return GGRT_ERR_UNSUPPORTED;
    }
    if ((long long)p->num_views * p->rays_per_view * (long long)(p->samples_per_ray > 0 ? p->samples_per_ray : 1) > 0x7fffffffLL) {
        set_error("adapter: more than 2^31 Gaussians");
        return GGRT_ERR_UNSUPPORTED;
    }
    return GGRT_OK;
}

int ggrt_adapter_forward(const GgrtAdapterParams* params, const float* extrinsics, const float* intrinsics,
                         const float* sh_rotation, const float* coordinates, const float* depths, const float* raw,
                         float* means, float* covariances, float* harmonics, float* scales_out,
                         float* rotations_out, ggrt_stream_t stream) {
    GGRT_TRY(check_adapter_params(params));
    const long long G = (long long)params->num_views * params->rays_per_view * params->samples_per_ray;
    if (G > 0 && (!extrinsics || !intrinsics || !coordinates || !depths || !raw || !means || !covariances || !harmonics)) {
        set_error("adapter_forward: NULL buffer");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NvtxRange nvtx_("adapter_forward");
    launch_adapter_forward(*params, extrinsics, intrinsics, sh_rotation, coordinates, depths, raw, means, covariances,
                           harmonics, scales_out, rotations_out, s);
    return check_launch("adapter_forward", 0, s);
}

int ggrt_adapter_backward(const GgrtAdapterParams* params, const float* extrinsics, const float* intrinsics,
                          const float* sh_rotation, const float* coordinates, const float* depths, const float* raw,
                          const float* dL_dmeans, const float* dL_dcovariances, const float* dL_dharmonics,
                          float* dL_dcoordinates, float* dL_ddepths, float* dL_draw, ggrt_stream_t stream) {
    GGRT_TRY(check_adapter_params(params));
    const long long G = (long long)params->num_views * params->rays_per_view * params->samples_per_ray;
    if (G > 0 && (!extrinsics || !intrinsics || !coordinates || !depths || !raw || !dL_dcoordinates || !dL_ddepths || !dL_draw)) {
        set_error("adapter_backward: NULL buffer");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NvtxRange nvtx_("adapter_backward");
    launch_adapter_backward(*params, extrinsics, intrinsics, sh_rotation, coordinates, depths, raw, dL_dmeans,
                            dL_dcovariances, dL_dharmonics, dL_dcoordinates, dL_ddepths, dL_draw, s);
    return check_launch("adapter_backward", 0, s);
}

int ggrt_camera_setup(int32_t num_views, const float* extrinsics, const float* intrinsics, const float* near,
                      const float* far, int32_t scale_invariant, float* cameras_out, ggrt_stream_t stream) {
    if (num_views < 0 || (num_views > 0 && (!extrinsics || !intrinsics || !near || !far || !cameras_out))) {
        set_error("camera_setup: bad argument");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    NvtxRange nvtx_("camera_setup");
    launch_camera_setup(num_views, extrinsics, intrinsics, near, far, scale_invariant, cameras_out, s);
    return check_launch("camera_setup", 0, s);
}

int ggrt_raster_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, uint8_t* present,
                             ggrt_stream_t stream) {
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
        set_error("mark_visible: bad argument");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    launch_mark_visible(P, means3D, viewmatrix, present, s);
    return check_launch("mark_visible", 0, s);
}

int ggrt_raster_profile_enable(int32_t on) {
    g_prof.on = on != 0;
    for (int i = 0; i < GGRT_STAGE_COUNT; ++i) g_prof.used[i] = false;
    return GGRT_OK;
}

int ggrt_raster_profile_read(float* ms_out) {
    if (!ms_out) {
        set_error("profile_read: NULL output");
        return GGRT_ERR_INVALID_ARGUMENT;
    }
    for (int i = 0; i < GGRT_STAGE_COUNT; ++i) {
        ms_out[i] = 0.f;
        if (!g_prof.made || !g_prof.used[i]) continue;
        if (cudaEventSynchronize(g_prof.b[i]) != cudaSuccess ||
            cudaEventElapsedTime(&ms_out[i], g_prof.a[i], g_prof.b[i]) != cudaSuccess) {
            cudaGetLastError();
            set_error("profile_read: event query failed for stage %d", i);
            return GGRT_ERR_CUDA;
        }
    }
    return GGRT_OK;
}

const char* ggrt_raster_stage_name(int32_t stage) {
    return (stage >= 0 && stage < GGRT_STAGE_COUNT) ? kStageNames[stage] : "?";
}

}  // extern "C"
