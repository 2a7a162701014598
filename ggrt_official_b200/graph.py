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
