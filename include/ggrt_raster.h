/*
 * ggrt_raster.h -- C ABI of libggrt_raster.so, the B200 (sm_100a) differentiable
 * 3D-Gaussian tile rasterizer that replaces the external `diff_gaussian_rasterization`
 * CUDA package on GGRt's render path.
 *
 * Reference interface replaced (the reference vendors no native code; these are the
 * extension entry points its Python front binds, reached from
 * /root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py:6-9 (import) and
 * :101-125 (the live call)):
 *
 *   _C.rasterize_gaussians(...)           -> ggrt_raster_forward_prepare + ggrt_raster_forward_render
 *   _C.rasterize_gaussians_backward(...)  -> ggrt_raster_backward
 *   _C.mark_visible(...)                  -> ggrt_raster_mark_visible
 *
 * Plain C: raw device pointers, sizes and a CUDA stream handle.  No torch / pybind
 * types cross this boundary.  The library never allocates or frees device memory and
 * keeps no data between calls; the caller owns every buffer (sizes from the
 * ggrt_raster_*_bytes functions).  The only internal objects are one side stream and four
 * events per (host thread, device): forward_prepare forks the SH colour kernel (and the zeroing
 * of the backward's scratch, GgrtRasterSettings.zero_scratch) onto it so that it overlaps the
 * binning kernels, forward_render joins it in front of the render kernel; ggrt_raster_backward
 * runs the dL/dsh writer on it beside the per-Gaussian kernel and joins before it returns
 * (GGRT_RASTER_OVERLAP=0 in the environment keeps every kernel on the caller's stream).
 * Consequently the inputs of forward_prepare must stay valid and unmodified until the matching
 * forward_render has been enqueued.  All float data is float32, all pointers are device
 * pointers unless the name ends in `_host`.  Every function returns 0 on success or a
 * negative GGRT_ERR_* code; ggrt_raster_last_error() returns a thread-local message.
 *
 * The forward is split in two because the number of tile-Gaussian pairs N is data
 * dependent and the caller owns the N-sized buffer: `prepare` culls/projects every
 * Gaussian, evaluates SH colours, counts pairs per tile and asynchronously copies
 * {N, max pairs in one tile} to pinned host memory; the caller waits for that copy
 * (colour evaluation keeps the GPU busy meanwhile), allocates
 * ggrt_raster_binning_bytes(N) and calls `render`.
 */
#ifndef GGRT_RASTER_H
#define GGRT_RASTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGRT_RASTER_ABI_VERSION 11
#define GGRT_RASTER_TILE 16 /* tile edge in pixels (tile ids are part of the contract) */
#define GGRT_RASTER_SUBS 16 /* pair counters per tile (contention spreading) */
#define GGRT_RASTER_MAX_MERGE_VIEWS 16 /* views per ggrt_raster_sh_gradient_merge call */

#define GGRT_OK 0
#define GGRT_ERR_INVALID_ARGUMENT (-1)
#define GGRT_ERR_CUDA (-2)
#define GGRT_ERR_UNSUPPORTED (-3)

typedef void* ggrt_stream_t; /* a cudaStream_t */

/* Mirrors GaussianRasterizationSettings as built at cuda_splatting.py:101-113. */
typedef struct GgrtRasterSettings {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    float scale_modifier; /* only meaningful with scales/rotations (unsupported: GGRt passes cov3D_precomp) */
    int32_t sh_degree;    /* 0..4 */
    int32_t prefiltered;  /* accepted, ignored (as upstream when cov3D is precomputed) */
    int32_t debug;        /* non-zero: synchronise and check for errors after every kernel */
    const float* viewmatrix; /* [4,4] device, row-major as passed: transposed world->camera */
    const float* projmatrix; /* [4,4] device: viewmatrix @ projection^T */
    const float* campos;     /* [3]   device */
    const float* bg;         /* [3]   device */
    /* Extensions for the sync-free render glue; zero / NULL = the upstream behaviour. */
    const float* device_params; /* [3] device: {tanfovx, tanfovy, scene_scale}.  When non-NULL these replace the host
                                   fields tanfovx / tanfovy above and GgrtRasterInputLayout.scene_scale, so that a caller
                                   whose camera lives on the device (ggrt_camera_setup) never reads it back -- the
                                   reference glue synchronises twice per view for them, cuda_splatting.py:104-105 */
    int32_t aux_mode;           /* 0: the aux channel is the `aux` argument (or the view depth).  1: GGRt's depth pass
                                   (cuda_splatting.py:240-268 with mode "depth"): aux = max(0, C0 * z + 0.5), z the
                                   camera-space depth of the UNSCALED scene, computed in the kernels; its gradient
                                   (dL_dout_aux) flows into dL_dmeans3D, dL_daux stays NULL */
    int32_t reserved;
    float* zero_scratch;        /* NULL, or the [P,12] grad_scratch buffer the caller will pass to ggrt_raster_backward for
                                   this frame: forward_prepare then zeroes it on the library's side stream, beside the
                                   colour kernel and under the binning kernels, and a backward that is given the same
                                   pointer in both places skips its own memset (which otherwise sits between the two
                                   render kernels on the critical path).  The caller must not touch the buffer between
                                   the two calls and must clear this field for a second backward of the same frame */
} GgrtRasterSettings;

/*
 * Camera set-up of GGRt's render glue for n views in ONE kernel and without a host read-back: replaces the ~60
 * small PyTorch operations and the two .item() synchronisations per view of cuda_splatting.py:64-89,104-105
 * (scale-invariant rescale by 1/near, get_fov, get_projection_matrix -- including its quirk of taking the
 * intrinsics of view 0 for every view, :39-42 -- the inverse of the camera-to-world matrix and the full projection).
 * extrinsics [n,4,4] camera-to-world, intrinsics [n,3,3] normalised, near / far [n]; all device float32.
 * Writes GGRT_CAMERA_FLOATS floats per view: [0,16) viewmatrix, [16,32) projmatrix (both as GgrtRasterSettings wants
 * them), [32,35) campos, [35,38) {tanfovx, tanfovy, scene_scale} = device_params, [38] near, [39] far after the rescale.
 */
#define GGRT_CAMERA_FLOATS 48
int ggrt_camera_setup(int32_t num_views, const float* extrinsics, const float* intrinsics, const float* near,
                      const float* far, int32_t scale_invariant, float* cameras_out, ggrt_stream_t stream);

/*
 * Optional description of how the caller stores the Gaussians, so that the copies GGRt's glue makes before
 * every rasterization can be skipped (cuda_splatting.py:66-77,116,124: scale-invariant rescale of means and
 * covariances, [g,3,n] -> [g,n,3] SH permute, upper-triangle gather).  NULL = the upstream layout.
 * Gradients are returned in the same layout and with respect to the UNSCALED inputs.
 */
typedef struct GgrtRasterInputLayout {
    float scene_scale;        /* means are multiplied by s, covariances by s*s before use (1.0 = none) */
    int32_t cov_full3x3;      /* 0: cov3D_precomp is [P,6]; 1: [P,3,3] row-major (upper triangle is read; the
                                 gradient is written to the upper triangle, zeros below) */
    int32_t sh_channel_major; /* 0: shs is [P,K,3]; 1: [P,3,K] */
} GgrtRasterInputLayout;

/*
 * Optional destinations of the compact colour gradients of ggrt_raster_backward (view-sharded multi-GPU
 * path, "push" model): instead of one local dL_dcolors [P,3], the kernel writes the rows -- staged in shared
 * memory and stored as coalesced 16-byte vectors -- to `count` buffers of [P+1,3] floats each, row P receiving
 * the view's campos.  The pointers may address peer-GPU memory (stores over NVLink are posted, so the remote
 * copies cost the kernel no latency), or, with multimem = 1, ptr[0] is an NVLS multicast address and ONE
 * multimem.st per vector lands in every GPU's buffer.  Each pointer must be 16-byte aligned.
 *
 * Step signalling (epoch != NULL; the multi-GPU exchange without barrier kernels and without the host): `epoch` is a
 * LOCAL device word counting completed pushes.  The kernel reads e = *epoch, pushes into half (e & 1) of every sink
 * (ptr[i] + (e & 1) * parity_stride floats: the receivers may still be reading the other half), and when its last
 * CTA has finished -- all stores fenced at system scope -- stores e + 1 to *epoch and adds 1 to every arrival
 * counter in `arrive` (release, system scope; arrive[0] is a multicast address when multimem).  A receiver that has
 * seen world * (e + 1) arrivals on its copy of the counter may read half (e & 1) of its sinks and, on every rank,
 * the plain outputs dL_dmeans3D / dL_dcov3D / dL_dopacity of this call (see ggrt_raster_sh_gradient_merge_signalled,
 * ggrt_raster_nvls_allreduce_signalled).  done_counter: a LOCAL device word, zero before the first launch.
 */
typedef struct GgrtRasterGradSinks {
    int32_t count;    /* 1..GGRT_RASTER_MAX_MERGE_VIEWS */
    int32_t multimem; /* 0: plain stores to every ptr[i]; 1: multimem.st to ptr[0] (count must be 1) */
    float* ptr[GGRT_RASTER_MAX_MERGE_VIEWS];
    uint32_t* epoch;        /* NULL: no signalling (the fields below are ignored) */
    uint32_t* done_counter;
    int64_t parity_stride;  /* floats */
    int32_t arrive_count;   /* 1 with multimem, else one counter per receiving GPU */
    int32_t reserved;
    uint32_t* arrive[GGRT_RASTER_MAX_MERGE_VIEWS];
} GgrtRasterGradSinks;

/* Byte offsets of the sub-arrays inside the caller-owned buffers (for tests / tools). */
typedef struct GgrtRasterLayout {
    /* geometry buffer, per Gaussian */
    size_t geom_rec0;   /* float4[P]  {pix_x, pix_y, cull threshold tau, view_depth} */
    size_t geom_rec1;   /* float4[P]  {conic_A, conic_B, conic_C, opacity} */
    size_t geom_rec2;   /* float4[P]  {r, g, b, aux} (aux = caller's 4th channel, default view_depth) */
    size_t geom_rect;   /* uint16x4[P] {tile_x0, tile_y0, tile_x1, tile_y1} */
    size_t geom_tiles;  /* uint32[P]  tiles touched */
    size_t geom_flags;  /* uint8[P]   bit c: colour channel c was clamped at 0 */
    size_t geom_ranks;  /* uint32x4[P] Gaussians touching <= 4 tiles: rank of each of their pairs inside its (tile, sub-counter)
                           segment, returned by the geometry kernel's counting atomics and consumed by the emit kernel */
    size_t geom_jac;    /* float[9][P] (SH inputs of degree > 0 only) plane 3c + j: d(colour channel c)/d(scaled mean j) through
                           the view direction, stored by the colour kernel while the SH row is on chip; the per-Gaussian
                           backward kernel multiplies it with dL/dcolour instead of reading the SH table a second time */
    size_t geom_bytes;
    /* image buffer, per tile / per pixel */
    size_t img_counts;  /* uint32[T*32] pairs per (tile, sub-counter); sub-counter = gaussian idx % 16 (+ 16 for Gaussians that
                           touch more than 4 tiles) */
    size_t img_partials;/* uint64[ceil(T/256)] per-scan-block {flag, max, total} (directly after img_counts, zeroed with it) */
    size_t img_cursor;  /* uint32[T*32] exclusive scan of img_counts (per (tile, sub-counter) segment starts); consumed as
                           allocation cursors by the emit kernel */
    size_t img_starts;  /* uint32[T+1]  exclusive scan of the per-tile totals; tile t owns [starts[t], starts[t+1]); [T] == N */
    size_t img_header;  /* uint32[4]   {N, max pairs in a tile, 0, 0} */
    size_t img_final_T; /* float[H*W]  */
    size_t img_ncontrib;/* uint32[H*W] */
    size_t img_bytes;
    /* binning buffer, per pair */
    size_t bin_keys;    /* uint64[N]  (depth_bits << 32 | gaussian_idx), sorted ascending inside each tile segment */
    size_t bin_points;  /* uint32[N]  gaussian idx, tile-major then depth then idx */
    size_t bin_masks;   /* uint8[N]   per list entry: bit w set if the Gaussian's {alpha >= 1/255} ellipse reaches warp pixel
                           block w of its tile (8x4 pixels; w & 1 = column, w >> 1 = row); written by the forward render
                           kernel for the entries it consumed, read by the backward render kernel */
    size_t bin_bytes;
} GgrtRasterLayout;

int ggrt_raster_abi_version(void);
const char* ggrt_raster_last_error(void);

/* Buffer sizes / layout.  num_rendered may be 0. */
int ggrt_raster_layout(int32_t P, int32_t image_height, int32_t image_width, int64_t num_rendered,
                       GgrtRasterLayout* out);
size_t ggrt_raster_geom_bytes(int32_t P);
size_t ggrt_raster_image_bytes(int32_t image_height, int32_t image_width);
size_t ggrt_raster_binning_bytes(int64_t num_rendered);

/*
 * Forward, phase 1.  Exactly one of shs [P,K,3] / colors_precomp [P,3] is non-NULL
 * (K = (sh_degree+1)^2).  cov3D_precomp is [P,6] = (xx,xy,xz,yy,yz,zz).  opacities [P].
 * aux [P] (may be NULL) is an extra per-Gaussian scalar that is alpha-blended with the same
 * weights as the colour into out_depth (NULL: the view-space depth is used).  GGRt's depth pass
 * (cuda_splatting.py:227-269) is such a channel, so colour and depth can share one rasterization.
 * Writes radii [P] (int32; 0 = culled), fills geom_buffer and the tile tables of
 * image_buffer, and the tile-scan kernel stores {N (low 32 bits), max pairs per tile, N >> 32, 0}
 * directly into counts_host (4 x uint32 of MAPPED pinned host memory -- cudaHostAlloc / torch
 * pin_memory under UVA; may be NULL if the caller reads img_header itself, which holds the same four
 * words).  A non-zero third word means N >= 2^32: such a frame cannot be rendered (pair lists are
 * indexed with 32 bits) and the caller must not call forward_render for it.  The values are valid once
 * work enqueued on the stream after this call (e.g. an event recorded right after it) has completed.
 */
int ggrt_raster_forward_prepare(const GgrtRasterSettings* settings, const GgrtRasterInputLayout* layout, int32_t P,
                                const float* means3D,
                                const float* cov3D_precomp, const float* opacities, const float* shs,
                                const float* colors_precomp, const float* aux, int32_t* radii, void* geom_buffer,
                                void* image_buffer, uint32_t* counts_host, ggrt_stream_t stream);

/*
 * Forward, phase 2.  binning_buffer holds ggrt_raster_binning_bytes(num_rendered) bytes.
 * num_rendered is its CAPACITY in pairs: normally the N that `prepare` reported, but a caller
 * that does not want to wait for N may pass a guess (e.g. the previous frame's N plus slack) and
 * launch immediately -- every kernel stays inside the capacity -- then compare the reported N with
 * its guess afterwards and, if N was larger, call again with rescan = 1 and a large enough buffer
 * (rescan re-runs the tile scan, whose cursors the first attempt consumed).  max_tile_pairs is
 * only a hint that selects the sort variant; larger tiles are still sorted correctly.  backward
 * must be given the same num_rendered.
 * Writes out_color [3,H,W], out_depth [H,W] (sum of aux * alpha * T -- view depth unless aux
 * was given -- no background, no normalisation) and the per-pixel state needed by backward.
 */
int ggrt_raster_forward_render(const GgrtRasterSettings* settings, int32_t P, int64_t num_rendered,
                               uint32_t max_tile_pairs, int32_t rescan, const void* geom_buffer, void* binning_buffer,
                               void* image_buffer, float* out_color, float* out_depth, ggrt_stream_t stream);

/*
 * Orders `stream` after the colour kernel that the most recent forward_prepare of this host thread forked onto
 * the library's side stream.  Only needed by a caller that abandons a frame between prepare and render
 * (forward_render joins by itself): after this call, work enqueued on `stream` -- including the release of the
 * buffers prepare was given to a stream-ordered allocator -- cannot overtake that kernel.
 */
int ggrt_raster_join(ggrt_stream_t stream);

/*
 * Backward.  dL_dout_color [3,H,W]; dL_dout_aux [H,W] or NULL (gradient of out_depth; only
 * meaningful when aux was given to prepare).  grad_scratch is [P,12] float32 scratch (zeroed
 * by the callee -- here, or already by forward_prepare, see GgrtRasterSettings.zero_scratch).  Outputs, all overwritten: dL_dmeans2D [P,3] (gradient w.r.t. NDC
 * xy, z = 0), dL_dopacity [P], dL_dmeans3D [P,3], dL_dcov3D [P,6], either dL_dsh [P,K,3]
 * (when shs was given) or dL_dcolors [P,3] (the unused one is NULL), and dL_daux [P] (NULL
 * unless dL_dout_aux is given).  dL_dcamera (NULL, or 35 floats: dL/dviewmatrix [4,4] |
 * dL/dprojmatrix [4,4] | dL/dcampos [3], overwritten) is an opt-in extension: the reference
 * treats the camera as constant (poses are detached, train_ggrt_stable.py:106).
 *
 * Compact mode (view-sharded multi-GPU path): with shs given, pass dL_dsh = NULL and dL_dcolors
 * [P,3] instead.  dL_dcolors then receives the gradient w.r.t. the evaluated SH colour (zero for
 * culled Gaussians and for channels clamped at 0) and no SH gradient is written; dL_dmeans3D still
 * contains the view-direction term.  dL/dsh of one view is basis(dir) (x) dL_dcolors, so the
 * per-view [P,3] arrays can be exchanged between GPUs instead of the K times larger SH gradients
 * and summed into dL/dsh with ggrt_raster_sh_gradient_merge.  color_sinks (may be NULL; compact mode
 * only, then dL_dcolors may be NULL) redirects / replicates that [P,3] output, see GgrtRasterGradSinks.
 */
int ggrt_raster_backward(const GgrtRasterSettings* settings, const GgrtRasterInputLayout* layout, int32_t P,
                         int64_t num_rendered, const float* means3D,
                         const float* cov3D_precomp, const float* shs, const int32_t* radii, const void* geom_buffer,
                         const void* binning_buffer, const void* image_buffer, const float* dL_dout_color,
                         const float* dL_dout_aux, float* grad_scratch, float* dL_dmeans2D, float* dL_dopacity,
                         float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dcolors, float* dL_daux,
                         float* dL_dcamera, const GgrtRasterGradSinks* color_sinks, ggrt_stream_t stream);

/*
 * dL_dsh[i] = sum over views v of basis(normalize(scene_scale * means3D[i] - campos_v)) (x) drgb_v[i]
 * for num_views <= GGRT_RASTER_MAX_MERGE_VIEWS views of the same P Gaussians: the SH gradient of a
 * multi-view step rebuilt from the compact per-view colour gradients of ggrt_raster_backward.
 * drgb_views_host / campos_views_host are HOST arrays of num_views device pointers ([P,3] and [3]
 * floats each); the pointers may address peer-GPU memory (P2P / symmetric memory over NVLink) -- the
 * kernel then gathers over NVLink while it streams dL_dsh to HBM -- provided the data was published
 * by a cross-device barrier the caller enqueued on `stream` before this call.  layout (may be NULL)
 * gives scene_scale and the SH layout of dL_dsh ([P,K,3], or [P,3,K] when sh_channel_major).
 * dL_dsh is overwritten.
 */
int ggrt_raster_sh_gradient_merge(int32_t P, int32_t sh_degree, const GgrtRasterInputLayout* layout,
                                  const float* means3D, int32_t num_views, const float* const* drgb_views_host,
                                  const float* const* campos_views_host, float* dL_dsh, ggrt_stream_t stream);

/*
 * The same merge for the signalled exchange (GgrtRasterGradSinks.epoch): the views are the `world` slots of THIS
 * GPU's slot table -- view v at slots + half * parity_stride + v * slot_stride floats, each slot [P+1,3] with the
 * view's campos in row P -- which the peers' backward kernels push into.  The kernel itself waits (acquire, system
 * scope) until `arrive` (this GPU's copy of the arrival counter) has reached world * *epoch, then reads half
 * (*epoch - 1) & 1.  Enqueue it after this rank's ggrt_raster_backward of the same step on the same stream.
 */
int ggrt_raster_sh_gradient_merge_signalled(int32_t P, int32_t sh_degree, const GgrtRasterInputLayout* layout,
                                            const float* means3D, int32_t world, const float* slots,
                                            int64_t slot_stride, int64_t parity_stride, const uint32_t* epoch,
                                            const uint32_t* arrive, float* dL_dsh, ggrt_stream_t stream);

/*
 * In-place float32 sum over `world` GPUs of `count` floats (a multiple of 4) that every rank holds at the
 * same offset of an NVLS multicast mapping (`multicast_ptr`: the multicast address of a symmetric
 * allocation, e.g. torch.distributed._symmetric_memory's multicast_ptr).  Rank r reduces slice r
 * inside the NVSwitch (multimem.ld_reduce) and broadcasts the result to all ranks (multimem.st).
 * The caller brackets the call with cross-device barriers on `stream` (inputs complete on every
 * rank before; results visible on every rank after).
 */
int ggrt_raster_nvls_allreduce_f32(void* multicast_ptr, int64_t count, int32_t rank, int32_t world,
                                   ggrt_stream_t stream);

/*
 * ggrt_raster_nvls_allreduce_f32 for the signalled exchange, with both waits inside the kernel: it starts reducing
 * when `arrive_in` (local copy) has reached world * *epoch -- every rank's backward outputs are complete -- and
 * returns when `arrive_out` has reached the same value, i.e. when every rank has broadcast its slice (its last CTA
 * adds 1 to all copies of arrive_out through `arrive_out_multicast` after its own slice).  done_counter: a LOCAL
 * device word, zero before the first launch.  May run beside the merge kernel on another stream.
 */
int ggrt_raster_nvls_allreduce_signalled(void* multicast_ptr, int64_t count, int32_t rank, int32_t world,
                                         const uint32_t* epoch, const uint32_t* arrive_in, void* arrive_out_multicast,
                                         const uint32_t* arrive_out, uint32_t* done_counter, ggrt_stream_t stream);

/*
 * Cross-GPU barrier on `stream` in one single-thread kernel (no host involvement): adds 1 to every GPU's copy of a
 * 32-bit counter through its NVLS multicast address (multimem.red, release) and waits until the local copy
 * (`local_counter`, the same symmetric allocation) has reached `target` = world * epoch, epoch = 1, 2, ... counted
 * by the caller (acquire; wrap-around safe).  The counter must be zero on every rank before the first use.
 */
int ggrt_raster_nvls_barrier(void* multicast_counter, const void* local_counter, uint32_t target, ggrt_stream_t stream);

/*
 * Fused Gaussian construction -- pixelSplat's "Gaussian adapter" (SURVEY.md 8f row 3).  Replaces the ~30 PyTorch
 * kernels of /root/reference/ggrt/model/pixelsplat/encoder/common/gaussian_adapter.py:48-96 (+ gaussians.py:8-44,
 * ggrt/geometry/projection.py:74-114, ggrt/misc/sh_rotation.py:10-29) by one forward and one backward kernel
 * that write the rasterizer's inputs directly.  Flattened "b v r srf spp" indexing: view = (b, v); a ray = (view,
 * r, srf) owns samples_per_ray Gaussians g = ray * samples_per_ray + s which share the ray's raw features and
 * image coordinate.  K = (sh_degree + 1)^2.
 *   extrinsics  [V,4,4]  camera-to-world          intrinsics [V,3,3]  normalised
 *   sh_rotation [V,K,K]  per-view SH rotation; only the (2l+1)x(2l+1) diagonal blocks are read (the matrices
 *                        e3nn's wigner_D gives the reference's rotate_sh); NULL = identity
 *   coordinates [R,2]    raw [R, 7 + 3K] = scales(3) | quaternion xyzw(4) | sh [3,K]     (R = V * rays_per_view)
 *   depths      [G]      (G = R * samples_per_ray)
 * Outputs: means [G,3], covariances [G,3,3], harmonics [G,3,K]; scales_out [G,3] / rotations_out [G,4] may be
 * NULL (the reference only uses them for .ply export).  Opacities pass through untouched and are not an argument.
 */
typedef struct GgrtAdapterParams {
    int32_t num_views;
    int32_t rays_per_view;
    int32_t samples_per_ray;
    int32_t sh_degree;     /* 0..4 */
    int32_t image_height;
    int32_t image_width;
    float scale_min;       /* GaussianAdapterCfg.gaussian_scale_min / _max */
    float scale_max;
    float eps;             /* quaternion normalisation eps of GaussianAdapter.forward (1e-8) */
} GgrtAdapterParams;

int ggrt_adapter_forward(const GgrtAdapterParams* params, const float* extrinsics, const float* intrinsics,
                         const float* sh_rotation, const float* coordinates, const float* depths, const float* raw,
                         float* means, float* covariances, float* harmonics, float* scales_out,
                         float* rotations_out, ggrt_stream_t stream);

/*
 * Backward of ggrt_adapter_forward: dL_dmeans [G,3], dL_dcovariances [G,3,3], dL_dharmonics [G,3,K] (any may be
 * NULL = zero) -> dL_dcoordinates [R,2], dL_ddepths [G], dL_draw [R, 7 + 3K], all overwritten; gradients of the
 * features a ray's samples share are summed inside the kernel (warp shuffles, no atomics).  Cameras are
 * constants, as in the reference (poses are detached).
 */
int ggrt_adapter_backward(const GgrtAdapterParams* params, const float* extrinsics, const float* intrinsics,
                          const float* sh_rotation, const float* coordinates, const float* depths, const float* raw,
                          const float* dL_dmeans, const float* dL_dcovariances, const float* dL_dharmonics,
                          float* dL_dcoordinates, float* dL_ddepths, float* dL_draw, ggrt_stream_t stream);

/* Frustum test only (upstream markVisible): present[i] = view-space z > 0.2. */
int ggrt_raster_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, uint8_t* present,
                             ggrt_stream_t stream);

/*
 * Per-kernel timing (diagnostics; off by default).  When enabled for the calling thread,
 * every kernel launch of the calls above is bracketed by CUDA events on the caller's
 * stream.  ggrt_raster_profile_read waits for the events of the most recent
 * forward_prepare / forward_render / backward calls and writes their durations in
 * milliseconds, indexed by GGRT_STAGE_* (a stage that did not run reports 0).
 */
#define GGRT_STAGE_GEOMETRY 0
#define GGRT_STAGE_SCAN_TILES 1
#define GGRT_STAGE_COLOR 2
#define GGRT_STAGE_EMIT 3
#define GGRT_STAGE_SORT_TILES 4
#define GGRT_STAGE_RENDER_FORWARD 5
#define GGRT_STAGE_RENDER_BACKWARD 6
#define GGRT_STAGE_PREPROCESS_BACKWARD 7
#define GGRT_STAGE_COUNT 8
int ggrt_raster_profile_enable(int32_t on);
int ggrt_raster_profile_read(float* ms_out /* [GGRT_STAGE_COUNT], host */);
const char* ggrt_raster_stage_name(int32_t stage);

#ifdef __cplusplus
}
#endif
#endif /* GGRT_RASTER_H */
