"""One training step of the render path -- forward, backward and (multi-GPU) the gradient exchange -- captured
once into a CUDA graph and replayed with a single launch.

The step is a fixed sequence of ~10 short kernels (0.3-0.4 ms in total at BASELINE config 2) whose launch
parameters depend only on the shape (P, H, W) and on the capacity of the pair buffer, so after one eager,
exactly-sized forward everything can be frozen: buffers come from a `rasterizer.Workspace`, the speculative pair
capacity is fixed with slack, and {N, max pairs per tile} of every replay lands in the workspace's pinned slot,
where `check()` compares it with the capacity after the fact.  The host then issues ONE launch per step and never
waits for the device, so host jitter (Python, the allocator, a profiler thread, eight ranks sharing the CPUs) can no
longer stall the GPU -- which at N > 1, where every step contains cross-GPU waits, is what scaling efficiency is
made of.  Inputs are read from the tensors given at construction: update them in place between replays.

Call site replaced: the per-view body of /root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py:93-127
plus autograd's backward of it, for a caller that renders the same shape every step.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import rasterizer as R


class CapturedStep:
    def __init__(self, means3D, sh, colors_precomp, opacities, cov3D_precomp, rs, grad_color: torch.Tensor,
                 aux=None, grad_aux: Optional[torch.Tensor] = None, layout: Optional[dict] = None, exchange=None,
                 want_camera: bool = False, slack: float = 1.5):
        """`grad_color` [3,H,W] (and `grad_aux` [H,W]) are the buffers the replayed backward reads dL/dimage from --
        a training loop writes its loss gradient into them (e.g. by capturing its loss kernels into the same graph
        with `extra_between`) or, as the benchmark does, keeps them fixed.  `exchange`: a
        view_parallel.CompactGradientExchange whose run() replaces the plain backward (one view per GPU)."""
        self.args = (means3D, sh, colors_precomp, opacities, cov3D_precomp, rs)
        self.aux, self.layout, self.exchange, self.want_camera = aux, layout, exchange, want_camera
        self.grad_color, self.grad_aux = grad_color, grad_aux
        dev = means3D.device
        self.device = dev
        # one eager, exactly checked forward: creates the library's side stream, fills the capacity estimate
        st = R.forward_raw(*self.args, aux=aux, layout=layout, check="sync")
        self.capacity = int(st["N"] * slack) + 4096
        c = st["call"]
        self.ws = R.Workspace(dev, c.P, c.H, c.W, self.capacity)
        self.out = None
        self.replays = 0
        # warm-up on a side stream (as torch.cuda.graph requires), then capture
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=s):
            self.out = self._body()
        self.state = self._state

    def _body(self):
        st = R.forward_raw(*self.args, aux=self.aux, layout=self.layout, workspace=self.ws, check="none")
        self._state = st
        if self.exchange is not None:
            return self.exchange.run(st, self.grad_color, want_camera=self.want_camera)
        if self.out is None:
            return R.backward_raw(st, self.grad_color, grad_aux=self.grad_aux, want_camera=self.want_camera)
        return R.backward_raw(st, self.grad_color, out={k: v for k, v in self.out.items() if v is not None and k != "dcamera"},
                              grad_aux=self.grad_aux, want_camera=self.want_camera)

    # outputs of the most recent replay (fixed tensors)
    @property
    def color(self):
        return self.ws.color

    @property
    def depth(self):
        return self.ws.depth

    @property
    def radii(self):
        return self.ws.radii

    @property
    def grads(self):
        return self.out

    def replay(self) -> None:
        """Enqueues the whole step on the current stream (one launch)."""
        self.graph.replay()
        self.replays += 1

    def check(self) -> tuple:
        """Waits for the replays issued so far and verifies the LAST one fitted its pair buffer (each replay
        overwrites the slot); raises rasterizer.BinningOverflow otherwise.  Returns (N, max pairs per tile)."""
        torch.cuda.current_stream(self.device).synchronize()
        N, mx = self.ws.counts.read()
        if N > self.ws.capacity:
            raise R.BinningOverflow(f"the captured step needed {N} tile-Gaussian pairs, its buffer holds "
                                    f"{self.ws.capacity}: outputs invalid; build a new CapturedStep")
        return N, mx


class CapturedViews:
    """ALL target views of one decoder call -- the loop of /root/reference/ggrt/model/pixelsplat/decoder/
    cuda_splatting.py:93-127 (colour) fused with :227-269 (depth) plus autograd's backward of it -- as ONE CUDA-graph
    launch on one GPU.

    What the graph holds for V views of the same Gaussians (pixelSplat layout: means [g,3], covariances [g,3,3],
    harmonics [g,3,K]):
      * one camera kernel for all views (`ggrt_camera_setup`; the rasterizer reads tan(fov/2) / scene scale from it),
      * per view the forward chain and the backward in COMPACT mode: the per-Gaussian backward writes the view's
        masked colour gradient [g,3] (12 B) instead of its SH gradient [g,3,K] (12K B),
      * ONE `sh_gradient_merge` that rebuilds dL/dharmonics = sum_v basis(dir_v) (x) dL/drgb_v for all views -- the
        SH gradient is written once per call instead of once per view plus an accumulation pass per view (what
        autograd does with V separate rasterizations: 3 x 12K B per Gaussian and view),
      * one reduction of the small per-view gradients (means, covariances, opacities).
    Views are issued round-robin on `streams` CUDA streams inside the capture, so the latency-, issue- and HBM-bound
    kernels of different views overlap.  The host issues one launch per call and never waits.

    Inputs are read from the tensors given here (update them in place between replays); `grad_color` [V,3,H,W] and
    `grad_depth` [V,H,W] (optional: turns on GGRt's depth channel, aux_mode = 1) are the buffers the backward reads
    dL/dimage from.  Results after `replay()`: `color` [V,3,H,W], `depth` [V,H,W], `grads` = {dmeans [g,3],
    dcovariances [g,3,3], dharmonics [g,3,K], dopacities [g]} summed over the views."""

    def __init__(self, extrinsics, intrinsics, near, far, image_shape, background_color, means, covariances,
                 harmonics, opacities, grad_color: torch.Tensor, grad_depth: Optional[torch.Tensor] = None,
                 scale_invariant: bool = True, streams: int = 2, slack: float = 1.5):
        from math import isqrt

        from . import render as RD

        dev = means.device
        self.device = dev
        V = int(extrinsics.shape[0])
        H, W = image_shape
        self.V, self.H, self.W = V, int(H), int(W)
        from . import _cabi

        if not 1 <= V <= _cabi.MAX_MERGE_VIEWS:
            raise ValueError(f"CapturedViews takes 1..{_cabi.MAX_MERGE_VIEWS} views per call, got {V}")
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        self.cam_in = (f(extrinsics), f(intrinsics), f(near).reshape(-1), f(far).reshape(-1))
        self.scale_invariant = bool(scale_invariant)
        self.bg = f(background_color).reshape(V, 3)
        self.means, self.cov, self.harm, self.opac = f(means), f(covariances), f(harmonics), f(opacities).reshape(-1)
        P = int(self.means.shape[0])
        K = int(self.harm.shape[-1])
        self.P, self.K, self.deg = P, K, isqrt(K) - 1
        if tuple(self.cov.shape) != (P, 3, 3) or tuple(self.harm.shape) != (P, 3, K) or (self.deg + 1) ** 2 != K:
            raise ValueError("CapturedViews expects pixelSplat's layout: means [g,3], covariances [g,3,3], harmonics [g,3,K]")
        if tuple(grad_color.shape) != (V, 3, self.H, self.W) or not grad_color.is_contiguous():
            raise ValueError(f"grad_color must be a contiguous [{V},3,{self.H},{self.W}] tensor")
        if grad_depth is not None and (tuple(grad_depth.shape) != (V, self.H, self.W) or not grad_depth.is_contiguous()):
            raise ValueError(f"grad_depth must be a contiguous [{V},{self.H},{self.W}] tensor")
        self.grad_color, self.grad_depth = grad_color, grad_depth
        self.layout = dict(scene_scale=1.0, cov_full3x3=True, sh_channel_major=True)  # the scale is read on the device
        f32 = dict(dtype=torch.float32, device=dev)
        self.cams = torch.empty((V, RD._cabi_camera_floats()), **f32)
        slot = (3 * (P + 1) + 3) // 4 * 4  # [g,3] colour gradients + the row the kernel adds (campos), float4-padded
        self.slots = torch.zeros((V, slot), **f32)
        self.small_names = [("dmeans3D", (P, 3)), ("dcov3D", (P, 3, 3)), ("dopacity", (P, 1))]
        self.dmeans2D = torch.empty((V, P, 3), **f32)    # per view (screen-space gradients are not summed over views)
        total = sum(int(torch.Size(s).numel()) for _, s in self.small_names)
        self.small = torch.empty((V, total), **f32)      # per-view small gradients, one flat row per view
        self.small_sum = torch.empty(total, **f32)
        self.cam_unscaled = torch.empty((V, 3), **f32)   # campos / scene scale: direction of (s m - c) = direction of (m - c / s)
        self.dsh = torch.empty((P, 3, K), **f32)
        self.n_streams = max(1, min(int(streams), V))
        # exactly sized eager forwards (one per view) give the capacities; then warm-up and capture
        RD.camera_setup(*self.cam_in, self.scale_invariant, out=self.cams)
        self.ws = []
        for v in range(V):
            st = R.forward_raw(self.means, self.harm, None, self.opac, self.cov, self._settings(v), layout=self.layout,
                               check="sync")
            self.ws.append(R.Workspace(dev, P, self.H, self.W, int(st["N"] * slack) + 4096))
        self.replays = 0
        s = torch.cuda.Stream(device=dev)
        self.pool = [s] + [torch.cuda.Stream(device=dev) for _ in range(self.n_streams - 1)]
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=s):
            self._body()

    def _settings(self, v: int):
        c = self.cams[v]
        return R.GaussianRasterizationSettings(
            image_height=self.H, image_width=self.W, tanfovx=0.0, tanfovy=0.0, bg=self.bg[v], scale_modifier=1.0,
            viewmatrix=c[0:16], projmatrix=c[16:32], sh_degree=self.deg, campos=c[32:35], prefiltered=False,
            device_params=c[35:38], aux_mode=1 if self.grad_depth is not None else 0)

    def _small_views(self, row: torch.Tensor) -> dict:
        out, off = {}, 0
        for name, shape in self.small_names:
            n = int(torch.Size(shape).numel())
            out[name] = row[off: off + n].view(shape)
            off += n
        return out

    def _body(self):
        from . import render as RD

        main = torch.cuda.current_stream(self.device)
        RD.camera_setup(*self.cam_in, self.scale_invariant, out=self.cams)
        torch.div(self.cams[:, 32:35], self.cams[:, 37:38], out=self.cam_unscaled)
        pool = [main] + self.pool[1:]
        for st in pool[1:]:
            st.wait_stream(main)
        for v in range(self.V):
            with torch.cuda.stream(pool[v % len(pool)]):
                st = R.forward_raw(self.means, self.harm, None, self.opac, self.cov, self._settings(v),
                                   layout=self.layout, workspace=self.ws[v], check="none")
                R.backward_raw(st, self.grad_color[v], out=dict(self._small_views(self.small[v]), dmeans2D=self.dmeans2D[v]),
                               grad_aux=None if self.grad_depth is None else self.grad_depth[v],
                               color_sinks={"ptrs": [self.slots[v].data_ptr()]})
        for st in pool[1:]:
            main.wait_stream(st)
        R.sh_gradient_merge(self.means, self.deg, [self.slots[v, : 3 * self.P] for v in range(self.V)],
                            [self.cam_unscaled[v] for v in range(self.V)], out=self.dsh,
                            layout=dict(scene_scale=1.0, sh_channel_major=True))
        torch.sum(self.small, dim=0, out=self.small_sum)

    @property
    def color(self) -> torch.Tensor:
        return torch.stack([w.color for w in self.ws])

    @property
    def depth(self) -> torch.Tensor:
        return torch.stack([w.depth for w in self.ws])

    @property
    def grads(self) -> dict:
        g = self._small_views(self.small_sum)
        return dict(dmeans=g["dmeans3D"], dcovariances=g["dcov3D"], dopacities=g["dopacity"].reshape(-1),
                    dharmonics=self.dsh)

    def replay(self) -> None:
        """Enqueues the whole multi-view step on the current stream (one launch)."""
        self.graph.replay()
        self.replays += 1

    def check(self) -> list:
        """Waits for the replays issued so far and verifies every view of the LAST one fitted its pair buffer."""
        torch.cuda.current_stream(self.device).synchronize()
        res = []
        for v, w in enumerate(self.ws):
            N, mx = w.counts.read()
            if N > w.capacity:
                raise R.BinningOverflow(f"view {v} of the captured call needed {N} tile-Gaussian pairs, its buffer holds "
                                        f"{w.capacity}: outputs invalid; build a new CapturedViews")
            res.append((N, mx))
        return res
