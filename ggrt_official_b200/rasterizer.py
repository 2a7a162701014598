"""Python front of the B200 rasterizer: the `diff_gaussian_rasterization` API surface.

Drop-in for the external package GGRt imports at
/root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py:6-9 and calls at :101-125:
same `GaussianRasterizationSettings` fields (with `debug` optional, because `render_cuda`
omits it, :101-113), same `GaussianRasterizer(settings)(means3D=..., means2D=..., ...)`
keywords, same argument validation, and a 3-tuple result `(color, radii, depth)` (the live
call site unpacks three values, :118).

Host code is PyTorch (allocation, streams, autograd); all compute is in libggrt_raster.so
behind the C ABI of include/ggrt_raster.h.  There is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _cabi


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool = False
    # extensions GGRt never passes (sync-free render glue, render.render_views_device):
    device_params: Optional[torch.Tensor] = None  # [3] CUDA {tanfovx, tanfovy, scene_scale}: replaces the host floats
    aux_mode: int = 0  # 1: third output = GGRt's depth pass, max(0, C0 z + 0.5) blended, differentiable w.r.t. means3D


def _debug_enabled(rs) -> bool:
    """`debug=True` in the settings, or GGRT_RASTER_DEBUG=1 in the environment: synchronise and check after every
    kernel, dump the inputs on failure (as upstream's debug mode)."""
    return bool(getattr(rs, "debug", False)) or os.environ.get("GGRT_RASTER_DEBUG") == "1"


_tls = threading.local()

# (device, P, H, W) -> (N, max pairs per tile) estimate for the next forward of that shape (see forward_raw): a
# slowly decaying high-water mark of the pair counts seen, so that a loop alternating between sparse and dense frames
# of one shape (other scenes, other crops) does not overflow its speculative pair buffer on every dense frame
_capacity_cache: dict = {}
SPECULATIVE_BINNING = True
CAPACITY_SLACK = 1.25
CAPACITY_DECAY = 0.97  # per forward; ~100 sparse frames forget a 20x denser one


def _remember_counts(key, N: int, max_pairs: int) -> None:
    prev = _capacity_cache.get(key)
    if prev is not None:
        N = max(int(N), int(prev[0] * CAPACITY_DECAY))
        max_pairs = max(int(max_pairs), int(prev[1] * CAPACITY_DECAY))
    _capacity_cache[key] = (int(N), int(max_pairs))
# how GaussianRasterizer (the autograd path) verifies the speculative pair buffer: "sync" (exact: one host wait per
# forward, like the reference, whose caller synchronises twice per view anyway) or "lazy" (no host wait; an overflow
# surfaces as BinningOverflow in backward / the next forward).  GGRT_RASTER_CHECK=lazy selects the latter.
AUTOGRAD_CHECK = "lazy" if os.environ.get("GGRT_RASTER_CHECK", "sync").lower() == "lazy" else "sync"


class BinningOverflow(RuntimeError):
    """A forward that ran with a deferred check (`check="lazy"` / a captured step) needed more tile-Gaussian
    pairs than its pair buffer held: that frame's outputs are invalid.  The capacity has been raised; redo it."""


class _Counts:
    """{N, max pairs per tile, N >> 32, 0}: 16 bytes of mapped pinned host memory the tile-scan kernel stores into."""

    def __init__(self):
        self.buf = torch.zeros(2, dtype=torch.int64).pin_memory()
        self.ptr = C.c_void_p(self.buf.data_ptr())

    def read(self):
        lo, hi = int(self.buf[0].item()), int(self.buf[1].item())
        if hi & 0xFFFFFFFF:
            raise RuntimeError("libggrt_raster forward failed (code -3): more than 2^32 - 1 tile-Gaussian pairs")
        return lo & 0xFFFFFFFF, (lo >> 32) & 0xFFFFFFFF


def _pinned_counts() -> "_Counts":
    """A ring of pinned slots per thread: with deferred checks several forwards can be in flight, and each needs
    the slot its scan kernel writes to stay untouched until it has been read."""
    ring = getattr(_tls, "ring", None)
    if ring is None:
        ring = _tls.ring = [[_Counts() for _ in range(8)], 0]
    ring[1] = (ring[1] + 1) % len(ring[0])
    return ring[0][ring[1]]


class Workspace:
    """Every buffer one rasterization of a fixed shape needs, allocated once: state (geometry, image tables, pair
    lists for `capacity` pairs), outputs (radii, colour, depth), the backward scratch and a private pinned slot for
    {N, max}.  A forward given a workspace performs no allocation, which makes the whole forward + backward a fixed
    launch sequence over fixed addresses -- what a CUDA graph needs (graph.CapturedStep) -- and spares the
    per-call allocator traffic in training loops that render the same shape every step."""

    def __init__(self, device, P: int, H: int, W: int, capacity: int):
        L = _cabi.lib()
        self.device = torch.device(device)
        self.P, self.H, self.W = int(P), int(H), int(W)
        u8 = dict(dtype=torch.uint8, device=self.device)
        self.radii = torch.empty(self.P, dtype=torch.int32, device=self.device)
        self.geom = torch.empty(L.ggrt_raster_geom_bytes(self.P), **u8)
        self.img = torch.empty(L.ggrt_raster_image_bytes(self.H, self.W), **u8)
        self.color = torch.empty((3, self.H, self.W), dtype=torch.float32, device=self.device)
        self.depth = torch.empty((self.H, self.W), dtype=torch.float32, device=self.device)
        self.scratch = torch.empty((self.P, 12), dtype=torch.float32, device=self.device)
        self.counts = _Counts()
        self.capacity = 0
        self.binning = None
        self.grow(capacity)

    def grow(self, capacity: int) -> None:
        if capacity > self.capacity:
            self.capacity = int(capacity)
            self.binning = torch.empty(_cabi.lib().ggrt_raster_binning_bytes(self.capacity), dtype=torch.uint8,
                                       device=self.device)


# forwards whose {N <= capacity} check was deferred: (event, counts slot, capacity, cache key)
def _pending() -> list:
    p = getattr(_tls, "pending", None)
    if p is None:
        p = _tls.pending = []
    return p


def check_pending(block: bool = True) -> None:
    """Verifies the deferred checks of earlier `check="lazy"` forwards of this thread (waiting for them if `block`);
    raises BinningOverflow if one of them overflowed its pair buffer."""
    p = _pending()
    bad = None
    while p:
        ev, counts, cap, key = p[0]
        if not block and len(p) < 6 and not ev.query():  # never more than 6 frames in flight: the pinned ring has 8 slots
            break
        ev.synchronize()
        p.pop(0)
        N, mx = counts.read()
        if key is not None:
            _remember_counts(key, N, mx)
        if N > cap:
            bad = (N, cap)
    if bad:
        raise BinningOverflow(f"a frame needed {bad[0]} tile-Gaussian pairs but its buffer held {bad[1]}; its outputs "
                              "are invalid -- the capacity estimate has been raised, redo the step")


def _f32c(t: Optional[torch.Tensor], name: str, device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the rasterizer has no CPU path")
    if t.device != device:
        raise RuntimeError(f"{name} is on {t.device}, expected {device}")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class _Call:
    """Validated, contiguous inputs + the C settings struct for one rasterization."""

    def __init__(self, means3D, sh, colors_precomp, opacities, cov3D_precomp, rs: GaussianRasterizationSettings,
                 aux=None, layout: Optional[dict] = None):
        if not means3D.is_cuda:
            raise RuntimeError("means3D must be a CUDA tensor: the rasterizer has no CPU path")
        dev = means3D.device
        self.device = dev
        self.means3D = _f32c(means3D, "means3D", dev)
        self.P = int(self.means3D.shape[0])
        self.sh = _f32c(sh, "shs", dev)
        self.colors = _f32c(colors_precomp, "colors_precomp", dev)
        self.opacities = _f32c(opacities, "opacities", dev).reshape(-1)
        self.cov3D = _f32c(cov3D_precomp, "cov3D_precomp", dev)
        self.aux = _f32c(aux, "aux_precomp", dev)
        if self.aux is not None:
            self.aux = self.aux.reshape(-1)
            if self.aux.numel() != self.P:
                raise ValueError(f"aux_precomp must have P={self.P} elements, got {self.aux.numel()}")
        self.H, self.W = int(rs.image_height), int(rs.image_width)
        self.deg = int(rs.sh_degree)
        K = (self.deg + 1) ** 2
        lay = layout or {}
        self.scale = float(lay.get("scene_scale", 1.0))
        self.cov9 = bool(lay.get("cov_full3x3", False))
        self.sh_cmajor = bool(lay.get("sh_channel_major", False))
        self.layout = None
        if layout:
            self.layout = _cabi.InputLayout(self.scale, int(self.cov9), int(self.sh_cmajor))
        if self.means3D.dim() != 2 or self.means3D.shape[1] != 3:
            raise ValueError(f"means3D must be [P,3], got {tuple(self.means3D.shape)}")
        if tuple(self.cov3D.shape) != ((self.P, 3, 3) if self.cov9 else (self.P, 6)):
            raise ValueError(f"cov3D_precomp must be {'[P,3,3]' if self.cov9 else '[P,6]'}, got {tuple(self.cov3D.shape)}")
        if self.opacities.numel() != self.P:
            raise ValueError(f"opacities must have P={self.P} elements, got {self.opacities.numel()}")
        if self.sh is not None and self.sh_cmajor:
            if tuple(self.sh.shape) != (self.P, 3, K):
                raise ValueError(f"channel-major shs must be [P,3,{K}], got {tuple(self.sh.shape)}")
        elif self.sh is not None:
            if self.sh.dim() != 3 or self.sh.shape[0] != self.P or self.sh.shape[2] != 3 or self.sh.shape[1] < K:
                raise ValueError(f"shs must be [P,>={K},3] for sh_degree {self.deg}, got {tuple(self.sh.shape)}")
            if self.sh.shape[1] != K:  # upstream reads the first (deg+1)^2 coefficients of a wider table
                self.sh = self.sh[:, :K, :].contiguous()
        if self.colors is not None and self.colors.shape != (self.P, 3):
            raise ValueError(f"colors_precomp must be [P,3], got {tuple(self.colors.shape)}")
        # settings tensors: always made contiguous (campos arrives as a stride-4 column slice, :111)
        self.view = _f32c(rs.viewmatrix, "viewmatrix", dev).reshape(16)
        self.proj = _f32c(rs.projmatrix, "projmatrix", dev).reshape(16)
        self.campos = _f32c(rs.campos, "campos", dev).reshape(3)
        self.bg = _f32c(rs.bg, "bg", dev).reshape(3)
        s = _cabi.Settings()
        s.image_height, s.image_width = self.H, self.W
        self.device_params = _f32c(getattr(rs, "device_params", None), "device_params", dev)
        self.aux_mode = int(getattr(rs, "aux_mode", 0) or 0)
        if self.device_params is not None:
            if self.device_params.numel() != 3:
                raise ValueError("device_params must hold {tanfovx, tanfovy, scene_scale}")
            s.device_params = self.device_params.data_ptr()
            s.tanfovx = s.tanfovy = 0.0
        else:
            s.tanfovx, s.tanfovy = float(rs.tanfovx), float(rs.tanfovy)
        s.aux_mode = self.aux_mode
        if self.aux_mode and self.aux is not None:
            raise ValueError("aux_mode=1 computes the aux channel itself: do not pass aux_precomp as well")
        s.scale_modifier = float(rs.scale_modifier)
        s.sh_degree = self.deg
        s.prefiltered = int(bool(rs.prefiltered))
        s.debug = int(_debug_enabled(rs))
        s.viewmatrix, s.projmatrix = self.view.data_ptr(), self.proj.data_ptr()
        s.campos, s.bg = self.campos.data_ptr(), self.bg.data_ptr()
        self.settings = s


def forward_raw(means3D, sh, colors_precomp, opacities, cov3D_precomp, rs: GaussianRasterizationSettings,
                aux=None, layout: Optional[dict] = None, workspace: Optional[Workspace] = None,
                check: str = "sync", prezero_scratch: bool = False) -> dict:
    """Runs the forward through the C ABI and returns outputs plus the opaque state buffers.
    `aux` [P]: optional extra per-Gaussian channel blended into the third output instead of the view depth.
    `layout`: optional {scene_scale, cov_full3x3, sh_channel_major} (struct GgrtRasterInputLayout).
    `workspace`: preallocated buffers (no allocation in this call; outputs live in the workspace).
    `check`: how {N <= capacity of the pair buffer} is verified --
      "sync"  wait for N inside this call and redo the binning if a speculative buffer was too small (exact, default);
      "lazy"  do not wait: the host returns at once and can queue further work; the check happens at the next
              forward of this thread (or check_pending()) and raises BinningOverflow after the fact;
      "none"  no check is scheduled (a captured CUDA graph: the owner reads workspace.counts after replays).
    `prezero_scratch`: a backward of this frame will follow -- its [P,12] scratch is allocated now and zeroed on the
      library's side stream under the binning kernels (GgrtRasterSettings.zero_scratch) instead of between the two
      render kernels.  Always on with a workspace (which owns the scratch)."""
    if check not in ("sync", "lazy", "none"):
        raise ValueError(f"check must be 'sync', 'lazy' or 'none', got {check!r}")
    L = _cabi.lib()
    c = _Call(means3D, sh, colors_precomp, opacities, cov3D_precomp, rs, aux, layout)
    dev = c.device
    ws = workspace
    if ws is not None and (ws.device != dev or (ws.P, ws.H, ws.W) != (c.P, c.H, c.W)):
        raise ValueError(f"workspace is for P={ws.P}, {ws.H}x{ws.W} on {ws.device}; this call has P={c.P}, "
                         f"{c.H}x{c.W} on {dev}")
    if check != "none":
        check_pending(block=(check == "sync"))  # surfaces an overflow of an earlier lazy frame; updates the estimates
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        sp = C.c_void_p(stream.cuda_stream)
        u8 = dict(dtype=torch.uint8, device=dev)
        if ws is None:
            radii = torch.empty(c.P, dtype=torch.int32, device=dev)
            geom = torch.empty(L.ggrt_raster_geom_bytes(c.P), **u8)
            img = torch.empty(L.ggrt_raster_image_bytes(c.H, c.W), **u8)
            color = torch.empty((3, c.H, c.W), dtype=torch.float32, device=dev)
            depth = torch.empty((c.H, c.W), dtype=torch.float32, device=dev)
            counts = _pinned_counts()
        else:
            radii, geom, img, color, depth, counts = ws.radii, ws.geom, ws.img, ws.color, ws.depth, ws.counts
        scratch = None
        if ws is not None:
            scratch = ws.scratch
        elif prezero_scratch and c.P > 0:
            scratch = torch.empty((c.P, 12), dtype=torch.float32, device=dev)
        if scratch is not None:
            c.settings.zero_scratch = scratch.data_ptr()
        lay = C.byref(c.layout) if c.layout is not None else None
        _cabi.check(L.ggrt_raster_forward_prepare(C.byref(c.settings), lay, c.P, _ptr(c.means3D), _ptr(c.cov3D),
                                                  _ptr(c.opacities), _ptr(c.sh), _ptr(c.colors), _ptr(c.aux),
                                                  _ptr(radii), _ptr(geom), _ptr(img), counts.ptr, sp),
                    "forward_prepare")
        try:
            ev = None
            if check != "none":
                ev = torch.cuda.Event()
                ev.record(stream)  # after the scan kernel's store of {N, max pairs per tile}; colour evaluation follows it

            def read_counts():
                ev.synchronize()
                return counts.read()

            def render(capacity, max_hint, rescan):
                if ws is None:
                    buf = torch.empty(L.ggrt_raster_binning_bytes(capacity), **u8)
                else:
                    ws.grow(capacity)
                    buf, capacity = ws.binning, ws.capacity
                _cabi.check(L.ggrt_raster_forward_render(C.byref(c.settings), c.P, capacity, max_hint, int(rescan),
                                                         _ptr(geom), _ptr(buf), _ptr(img), _ptr(color), _ptr(depth), sp),
                            "forward_render")
                return buf, capacity

            # N sizes the caller-owned pair buffer.  Waiting for it costs a host round trip that the colour kernel is
            # too short to hide, so after the first call of a given shape the buffer is sized from the previous N plus
            # slack and the remaining kernels are launched at once; N is checked afterwards and the binning + render is
            # redone with the exact size in the rare case the guess was too small ("sync"), or the check is left to
            # the next call so that the host never waits for the device ("lazy").
            key = (dev.index, c.P, c.H, c.W)
            guess = _capacity_cache.get(key) if SPECULATIVE_BINNING else None
            if guess is None and check != "sync":
                if ws is None or ws.capacity == 0:
                    raise RuntimeError("check='lazy'/'none' needs a capacity estimate: run one check='sync' forward of "
                                       "this shape first (or pass a workspace with a capacity)")
                guess = (ws.capacity, 0)
            if guess is None:
                N, max_pairs = read_counts()
                binning, cap = render(N, max_pairs, False)
            else:
                want = int(guess[0] * CAPACITY_SLACK) + 1024 if ws is None or ws.capacity == 0 else ws.capacity
                hint = int(guess[1] * 1.05) + 4 if guess[1] else 0xFFFFFF  # unknown: the general sort tier
                binning, cap = render(want, hint, False)  # hint only: larger tiles still sort correctly
                if check == "sync":
                    N, max_pairs = read_counts()
                    if N > cap:
                        binning, cap = render(N, max_pairs, True)
                else:
                    N, max_pairs = guess
                    if check == "lazy":
                        _pending().append((ev, counts, cap, key))
            if check == "sync":
                _remember_counts(key, N, max_pairs)
        except Exception:
            # the colour kernel forked by `prepare` may still be running on the library's side stream: order the
            # caller's stream (and with it the release of geom / radii to the caching allocator) after it
            L.ggrt_raster_join(sp)
            raise
    return dict(call=c, color=color, depth=depth, radii=radii, geom=geom, img=img, binning=binning, N=N,
                capacity=cap, max_tile_pairs=max_pairs, workspace=ws, scratch=scratch)


def backward_raw(state: dict, grad_color: torch.Tensor, out: Optional[dict] = None,
                 grad_aux: Optional[torch.Tensor] = None, want_camera: bool = False, compact: bool = False,
                 color_sinks: Optional[dict] = None) -> dict:
    """Runs the backward through the C ABI.  `out` may supply preallocated, contiguous float32 output
    tensors (e.g. views into one gradient arena that is all-reduced across GPUs afterwards).
    `grad_aux` [H,W]: gradient of the third output (only when the forward was given `aux`).
    `want_camera`: also return out["dcamera"] = dL/d(viewmatrix [16] | projmatrix [16] | campos [3]) (opt-in
    extension; the reference treats the camera as constant).
    `compact` (SH inputs only): skip dL/dsh and return out["dcolors"] [P,3], the gradient w.r.t. the evaluated SH
    colour, instead -- dL/dsh of one view is basis(dir) (x) dcolors, which `sh_gradient_merge` rebuilds for a whole
    set of views after the (K times smaller) colour gradients have been exchanged between GPUs.
    `color_sinks` = {"ptrs": [device addresses], "multimem": bool} (implies compact): the kernel itself writes the
    [P,3] colour gradients plus a row P with the view's campos to each of the 16-byte aligned [P+1,3] buffers --
    peer-GPU memory, or one NVLS multicast address with multimem=True (struct GgrtRasterGradSinks); no local
    "dcolors" is returned.  Optional keys "epoch", "done", "parity_stride", "arrive" switch on the in-kernel step
    signalling of the same struct (view_parallel.CompactGradientExchange uses it)."""
    L = _cabi.lib()
    c: _Call = state["call"]
    sinks = None
    if color_sinks is not None:
        compact = True
        ptrs = [int(p) for p in color_sinks["ptrs"]]
        sinks = _cabi.GradSinks()
        sinks.count, sinks.multimem = len(ptrs), int(bool(color_sinks.get("multimem", False)))
        if not 1 <= len(ptrs) <= _cabi.MAX_MERGE_VIEWS:
            raise ValueError(f"color_sinks needs 1..{_cabi.MAX_MERGE_VIEWS} pointers, got {len(ptrs)}")
        for k, p in enumerate(ptrs):
            sinks.ptr[k] = p
        if color_sinks.get("epoch"):  # in-kernel step signalling (struct GgrtRasterGradSinks)
            arrive = [int(a) for a in color_sinks["arrive"]]
            sinks.epoch, sinks.done_counter = int(color_sinks["epoch"]), int(color_sinks["done"])
            sinks.parity_stride, sinks.arrive_count = int(color_sinks["parity_stride"]), len(arrive)
            for k, a in enumerate(arrive):
                sinks.arrive[k] = a
    if compact and c.sh is None:
        raise ValueError("compact=True needs SH inputs (colors_precomp already yields dcolors)")
    dev = c.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        sp = C.c_void_p(stream.cuda_stream)
        g = _f32c(grad_color, "grad_color", dev)
        ga = _f32c(grad_aux, "grad_aux", dev)
        if ga is not None and c.aux is None and not c.aux_mode:
            raise RuntimeError("the third output is only differentiable when aux_precomp was given")
        f32 = dict(dtype=torch.float32, device=dev)
        # the scratch the forward zeroed for this frame (settings.zero_scratch still names it), else a fresh one
        # that the library zeroes itself
        scratch = state.get("scratch") if c.settings.zero_scratch else None
        if scratch is None:
            ws = state.get("workspace")
            scratch = ws.scratch if ws is not None else torch.empty((c.P, 12), **f32)
            c.settings.zero_scratch = None
        given = out or {}

        def buf(name, shape):
            t = given.get(name)
            if t is None:
                return torch.empty(shape, **f32)
            if tuple(t.shape) != tuple(shape) or t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                raise ValueError(f"out[{name!r}] must be a contiguous float32 {tuple(shape)} tensor on {dev}")
            return t

        out = dict(
            dmeans2D=buf("dmeans2D", (c.P, 3)),
            dopacity=buf("dopacity", (c.P, 1)),
            dmeans3D=buf("dmeans3D", (c.P, 3)),
            dcov3D=buf("dcov3D", (c.P, 3, 3) if c.cov9 else (c.P, 6)),
            dsh=buf("dsh", tuple(c.sh.shape)) if c.sh is not None and not compact else None,
            dcolors=buf("dcolors", (c.P, 3)) if c.sh is None or (compact and sinks is None) else None,
            daux=buf("daux", (c.P,)) if ga is not None and not c.aux_mode else None,
            dcamera=torch.zeros(35, **f32) if want_camera else None,
        )
        lay = C.byref(c.layout) if c.layout is not None else None
        _cabi.check(L.ggrt_raster_backward(C.byref(c.settings), lay, c.P, state["capacity"], _ptr(c.means3D), _ptr(c.cov3D),
                                           _ptr(c.sh), _ptr(state["radii"]), _ptr(state["geom"]),
                                           _ptr(state["binning"]), _ptr(state["img"]), _ptr(g), _ptr(ga),
                                           _ptr(scratch), _ptr(out["dmeans2D"]), _ptr(out["dopacity"]),
                                           _ptr(out["dmeans3D"]), _ptr(out["dcov3D"]), _ptr(out["dsh"]),
                                           _ptr(out["dcolors"]), _ptr(out["daux"]), _ptr(out["dcamera"]),
                                           C.byref(sinks) if sinks is not None else None, sp),
                    "backward")
        c.settings.zero_scratch = None  # used up: a second backward of this frame has the library zero the scratch
    return out


def sh_gradient_merge(means3D: torch.Tensor, sh_degree: int, drgb_views, campos_views,
                      out: Optional[torch.Tensor] = None, layout: Optional[dict] = None) -> torch.Tensor:
    """dL/dsh [P,K,3] of a set of views of the same Gaussians from their compact colour gradients:
    sum_v basis(normalize(scene_scale * means3D - campos_v)) (x) drgb_v   (ggrt_raster_sh_gradient_merge).

    `drgb_views` / `campos_views`: per view either a float32 CUDA tensor ([P,3] / [3], contiguous) or an integer
    device address -- e.g. a peer GPU's symmetric-memory buffer, in which case the gather over NVLink happens
    inside the kernel; the caller is responsible for the cross-device barrier before this call."""
    L = _cabi.lib()
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor: the rasterizer has no CPU path")
    dev = means3D.device
    means = _f32c(means3D, "means3D", dev)
    P = int(means.shape[0])
    K = (int(sh_degree) + 1) ** 2
    V = len(drgb_views)
    if V != len(campos_views) or not 1 <= V <= _cabi.MAX_MERGE_VIEWS:
        raise ValueError(f"need 1..{_cabi.MAX_MERGE_VIEWS} views with one campos each, got {V} / {len(campos_views)}")
    lay = layout or {}
    cmajor = bool(lay.get("sh_channel_major", False))
    shape = (P, 3, K) if cmajor else (P, K, 3)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=dev)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise ValueError(f"out must be a contiguous float32 {shape} tensor on {dev}")
    keep = []

    def address(x, name, numel):
        if isinstance(x, int):
            return x
        x = _f32c(x, name, dev)
        if x.numel() != numel:
            raise ValueError(f"{name} must have {numel} elements, got {x.numel()}")
        keep.append(x)
        return x.data_ptr()

    drgb = (C.c_void_p * V)(*[address(x, "drgb_views[i]", 3 * P) for x in drgb_views])
    cams = (C.c_void_p * V)(*[address(x, "campos_views[i]", 3) for x in campos_views])
    clay = _cabi.InputLayout(float(lay.get("scene_scale", 1.0)), int(bool(lay.get("cov_full3x3", False))), int(cmajor))
    with torch.cuda.device(dev):
        sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(L.ggrt_raster_sh_gradient_merge(P, int(sh_degree), C.byref(clay), _ptr(means), V, drgb, cams,
                                                    _ptr(out), sp), "sh_gradient_merge")
    return out


def _dump(path: str, payload) -> None:
    try:
        torch.save(payload, path)
        print(f"\nAn error occurred in the rasterizer. Inputs were written to {path} for debugging.")
    except Exception:  # pragma: no cover - best effort
        pass


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, aux=None, layout=None, viewmatrix=None, projmatrix=None, campos=None):
        # viewmatrix / projmatrix / campos are the tensors of raster_settings again: passing them as explicit
        # inputs lets autograd deliver camera gradients when (and only when) they require grad
        try:
            st = forward_raw(means3D, sh, colors_precomp, opacities, cov3Ds_precomp, raster_settings, aux, layout,
                             check=AUTOGRAD_CHECK, prezero_scratch=any(ctx.needs_input_grad))
        except Exception:
            if _debug_enabled(raster_settings):
                _dump("snapshot_fw.dump", (means3D, sh, colors_precomp, opacities, cov3Ds_precomp, raster_settings))
            raise
        # keep only what backward needs; the outputs must not be referenced from ctx (that would be a
        # grad_fn <-> output reference cycle and the buffers would wait for the garbage collector)
        ctx.state = {k: st[k] for k in ("call", "radii", "geom", "img", "binning", "N", "capacity", "scratch")}
        ctx.raster_settings = raster_settings
        ctx.sh_shape = None if sh is None else tuple(sh.shape)
        ctx.opacity_shape = tuple(opacities.shape)
        ctx.has_aux = aux is not None or bool(getattr(raster_settings, "aux_mode", 0))
        ctx.aux_shape = None if aux is None else tuple(aux.shape)
        if ctx.has_aux:  # the third output blends the caller's channel and is differentiable
            ctx.mark_non_differentiable(st["radii"])
        else:
            ctx.mark_non_differentiable(st["radii"], st["depth"])
        ctx.set_materialize_grads(False)
        return st["color"], st["radii"], st["depth"]

    @staticmethod
    def backward(ctx, grad_color, _grad_radii=None, _grad_depth=None):
        st = ctx.state
        c: _Call = st["call"]
        grad_aux = _grad_depth if ctx.has_aux else None
        want_cam = any(ctx.needs_input_grad[11:14])
        if grad_color is None and grad_aux is None:
            return (None,) * 14
        if grad_color is None:
            grad_color = torch.zeros((3, c.H, c.W), dtype=torch.float32, device=c.device)
        if AUTOGRAD_CHECK == "lazy":
            check_pending(block=True)  # the forward of this frame has long finished: costs nothing, raises on overflow
        try:
            g = backward_raw(st, grad_color, grad_aux=grad_aux, want_camera=want_cam)
        except Exception:
            if _debug_enabled(ctx.raster_settings):
                _dump("snapshot_bw.dump", (c.means3D, c.sh, c.colors, c.opacities, c.cov3D, grad_color))
            raise
        dsh = g["dsh"]
        if dsh is not None and ctx.sh_shape != tuple(dsh.shape):  # wider SH table than (deg+1)^2: pad with zeros
            full = torch.zeros(ctx.sh_shape, dtype=dsh.dtype, device=dsh.device)
            full[:, : dsh.shape[1]] = dsh
            dsh = full
        daux = g.get("daux")
        if daux is not None and ctx.aux_shape is not None:
            daux = daux.reshape(ctx.aux_shape)
        dview = dproj = dcampos = None
        if want_cam:
            rs = ctx.raster_settings
            cam = g["dcamera"]
            dview = cam[:16].reshape(rs.viewmatrix.shape) if ctx.needs_input_grad[11] else None
            dproj = cam[16:32].reshape(rs.projmatrix.shape) if ctx.needs_input_grad[12] else None
            dcampos = cam[32:35].reshape(rs.campos.shape) if ctx.needs_input_grad[13] else None
        return (g["dmeans3D"], g["dmeans2D"] if ctx.needs_input_grad[1] else None, dsh, g["dcolors"],
                g["dopacity"].reshape(ctx.opacity_shape), None, None, g["dcov3D"], None, daux, None, dview, dproj,
                dcampos)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, aux_precomp=None, layout=None):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, aux_precomp, layout, raster_settings.viewmatrix,
                                     raster_settings.projmatrix, raster_settings.campos)


def covariance_from_scaling_rotation(scales: torch.Tensor, rotations: torch.Tensor, scale_modifier: float = 1.0):
    """[P,3] scales + [P,4] quaternions (r, x, y, z) -> [P,6] upper-triangular world covariances (xx, xy, xz, yy, yz, zz),
    as upstream's computeCov3D does inside its preprocess kernel."""
    if scales.dim() != 2 or scales.shape[1] != 3 or rotations.dim() != 2 or rotations.shape[1] != 4 or \
            scales.shape[0] != rotations.shape[0]:
        raise ValueError(f"scales must be [P,3] and rotations [P,4], got {tuple(scales.shape)} / {tuple(rotations.shape)}")
    r, x, y, z = rotations.float().unbind(dim=1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)
    M = R * (scales.float() * scale_modifier)[:, None, :]  # R S
    cov = M @ M.transpose(1, 2)                            # R S S^T R^T
    iu = torch.triu_indices(3, 3, device=cov.device)
    return cov[:, iu[0], iu[1]].contiguous()


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """Frustum test of upstream `markVisible` (unused by GGRt; kept for API completeness)."""
        L = _cabi.lib()
        rs = self.raster_settings
        with torch.no_grad():
            if not positions.is_cuda:
                raise RuntimeError("positions must be a CUDA tensor: the rasterizer has no CPU path")
            pos = _f32c(positions, "positions", positions.device)
            view = _f32c(rs.viewmatrix, "viewmatrix", positions.device)
            out = torch.empty(pos.shape[0], dtype=torch.uint8, device=pos.device)
            with torch.cuda.device(pos.device):
                sp = C.c_void_p(torch.cuda.current_stream(pos.device).cuda_stream)
                _cabi.check(L.ggrt_raster_mark_visible(pos.shape[0], _ptr(pos), _ptr(view), _ptr(out), sp),
                            "mark_visible")
        return out.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, aux_precomp=None, layout=None):
        """Same keywords as upstream, plus two extensions GGRt never passes: `aux_precomp` [P] or [P,1], an extra
        per-Gaussian scalar that is alpha-blended with the colour's weights into the third output (differentiable;
        without it the third output is the blended view-space depth, not differentiable), and `layout`, a dict
        {scene_scale, cov_full3x3, sh_channel_major} describing inputs stored as pixelSplat's `Gaussians` holds
        them, so the rescale / gather / permute copies of the reference glue can be skipped."""
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if cov3D_precomp is None:
            # GGRt always passes cov3D_precomp (cuda_splatting.py:124).  For callers of the stock 3DGS interface the
            # covariance is built here with differentiable PyTorch operations (upstream computeCov3D: Sigma = R S^2 R^T,
            # S = scale_modifier * diag(scales), R from the quaternion (r, x, y, z) as given -- not re-normalised) and
            # handed to the same kernels; gradients reach scales / rotations through autograd.
            cov3D_precomp = covariance_from_scaling_rotation(scales, rotations, float(self.raster_settings.scale_modifier))
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   self.raster_settings, aux_precomp, layout)
