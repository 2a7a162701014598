"""Multi-GPU scaling of the render path: independent target views are sharded across ranks.

Each target view is an independent rasterization of the same Gaussian set (the per-view
loop at /root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py:93-127 has no
cross-iteration dependence), so rank r renders views r, r+world, ... of the replicated
Gaussians with no data-path collective in the forward.  The only exchange step is the sum of
the per-view Gaussian gradients.  One process per GPU, torch.distributed (NCCL
over NVLink on the B200 box; gloo in the CPU tests).  Two implementations of the exchange:

* `GradientArena`: ONE all-reduce over a single contiguous arena [P, 3 + 6 + 1 + 3K]
  (340 B per Gaussian at SH degree 4) -- the baseline.
* `CompactGradientExchange`: dL/dsh of one view is the outer product basis(dir_v) (x) dL/drgb_v
  and every rank can evaluate basis(dir_v) itself, so only the [P,3] colour gradients are
  gathered (12 B per Gaussian and view) and the [P,10] rest is all-reduced; the summed SH
  gradient is rebuilt locally by `sh_gradient_merge`.  With transport "p2p" the buffers are
  symmetric memory: the merge kernel gathers straight from the peers over NVLink and the small
  arena is summed in the NVSwitch (multimem) -- no NCCL call on the data path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment of target views to ranks."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    return list(range(rank, num_views, world_size))


@dataclass
class GradientArena:
    """One flat float32 buffer holding every Gaussian gradient, plus named views into it."""

    flat: torch.Tensor
    views: Dict[str, torch.Tensor]

    @staticmethod
    def allocate(P: int, K: int, device, use_sh: bool = True) -> "GradientArena":
        names = [("dmeans3D", (P, 3)), ("dcov3D", (P, 6)), ("dopacity", (P, 1)),
                 ("dsh", (P, K, 3)) if use_sh else ("dcolors", (P, 3))]
        total = sum(int(torch.Size(s).numel()) for _, s in names)
        flat = torch.empty(total, dtype=torch.float32, device=device)
        views, off = {}, 0
        for name, shape in names:
            n = int(torch.Size(shape).numel())
            views[name] = flat[off: off + n].view(shape)
            off += n
        return GradientArena(flat, views)

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum over ranks, in place (float32: the 1e-3 relative gradient tolerance leaves no room for compression)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def all_reduce_gradients(tensors: Sequence[Optional[torch.Tensor]], group=None) -> None:
    """Sums `.grad` of the given leaf tensors over ranks with one collective (autograd-level path).
    Ranks that rendered no view contribute zeros."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    leaves = [t for t in tensors if t is not None]
    for t in leaves:
        if t.grad is None:
            t.grad = torch.zeros_like(t)
    flat = torch.cat([t.grad.reshape(-1) for t in leaves])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in leaves:
        n = t.grad.numel()
        t.grad.copy_(flat[off: off + n].view_as(t.grad))
        off += n


def render_views_sharded(render_fn, num_views: int, rank: Optional[int] = None, world_size: Optional[int] = None):
    """Calls `render_fn(view_index)` for the views owned by this rank; returns {view_index: result}."""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return {v: render_fn(v) for v in shard_views(num_views, rank, world_size)}


class CompactGradientExchange:
    """Sum of the per-view Gaussian gradients over ranks that moves 12 B + 2 x 40 B instead of 2 x 340 B per
    Gaussian (SH degree 4) across NVLink; see the module docstring.  Buffers are allocated once.

    run(state, grad_color) performs this rank's backward in compact mode, the exchange and the merge and
    returns {"dmeans3D", "dcov3D", "dopacity", "dsh"} summed over all ranks plus this view's "dmeans2D".
    `backward_fn` / `merge_fn` default to the CUDA entry points (`rasterizer.backward_raw`,
    `rasterizer.sh_gradient_merge`); the CPU tests inject oracle-backed ones to exercise the choreography on gloo.
    """

    def __init__(self, P: int, sh_degree: int, device, group=None, transport: str = "nccl",
                 layout: Optional[dict] = None, backward_fn: Optional[Callable] = None,
                 merge_fn: Optional[Callable] = None, barrier: str = "signal"):
        if transport not in ("nccl", "p2p"):
            raise ValueError(f"unknown transport {transport!r}")
        on = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.P, self.deg, self.K = int(P), int(sh_degree), (int(sh_degree) + 1) ** 2
        self.device = torch.device(device)
        self.layout = layout
        self.transport = transport if self.world > 1 else "nccl"
        if layout and layout.get("cov_full3x3"):
            raise NotImplementedError("CompactGradientExchange expects [P,6] covariances")
        from . import rasterizer as R

        self.backward_fn = backward_fn or R.backward_raw
        self.merge_fn = merge_fn or R.sh_gradient_merge
        f32 = dict(dtype=torch.float32, device=self.device)
        self.slot = (3 * (self.P + 1) + 3) // 4 * 4  # [P,3] colour gradients + a row with the view's campos (float4-padded)
        self.small_elems = (10 * self.P + 3) // 4 * 4  # dmeans3D | dcov3D | dopacity, padded to float4
        self.handles = None
        self.sinks = None
        self._side = None
        self.profile, self.marks = False, []
        self.barrier_kind = "torch"
        self.signalled = False
        if self.transport == "p2p":
            import torch.distributed._symmetric_memory as symm

            # "push" model: every rank holds the slots of ALL views in symmetric memory and each rank's backward
            # kernel stores its colour gradients into slot `rank` of every GPU (posted NVLink stores, or ONE
            # multimem.st per vector when the fabric has NVLS multicast), so the merge only reads local memory
            gname = (group or dist.group.WORLD).group_name
            probe = symm.empty(4, dtype=torch.int32, device=self.device)
            has_mc = bool(symm.rendezvous(probe, gname).multicast_ptr)
            # Synchronisation between the ranks: with NVLS multicast the kernels signal and wait themselves
            # ("signal": arrival counters bumped with multimem.red by the last CTA of the producing kernel, polled by
            # the first instruction of the consuming one -- no barrier kernels, no host, double-buffered slots);
            # otherwise torch's symmetric-memory barrier kernel ("torch") or the single-thread barrier kernel of this
            # library ("nvls") brackets the exchange.
            if barrier == "signal" and not has_mc:
                barrier = "torch"
            self.signalled = barrier == "signal"
            self.halves = 2 if self.signalled else 1
            self.slots = symm.empty(self.halves * self.world * self.slot, **f32)
            self.small = symm.empty(self.small_elems, **f32)
            self.handles = (symm.rendezvous(self.slots, gname), symm.rendezvous(self.small, gname))
            self.my_slot = self.slots[self.rank * self.slot: (self.rank + 1) * self.slot]
            hs = self.handles[0]
            off = 4 * self.rank * self.slot
            self.barrier_kind = barrier if hs.multicast_ptr else "torch"
            if self.barrier_kind in ("nvls", "signal"):
                self.bar = symm.empty(8, dtype=torch.int32, device=self.device)  # [0] push arrivals, [1] reduce arrivals
                self.bar.zero_()
                self.bar_handle = symm.rendezvous(self.bar, gname)
                self.ctrl = torch.zeros(4, dtype=torch.int32, device=self.device)  # [0] epoch, [1], [2] CTA counters
                torch.cuda.synchronize(self.device)
                hs.barrier(channel=0)  # every rank's counters are zero before anyone increments them
                torch.cuda.synchronize(self.device)
                self.epoch = 0
            if hs.multicast_ptr:
                self.sinks = {"ptrs": [int(hs.multicast_ptr) + off], "multimem": True}
            else:
                self.sinks = {"ptrs": [int(p) + off for p in hs.buffer_ptrs], "multimem": False}
            if self.signalled:
                self.sinks.update(epoch=self.ctrl.data_ptr(), done=self.ctrl.data_ptr() + 4,
                                  parity_stride=self.world * self.slot, arrive=[int(self.bar_handle.multicast_ptr)])
        else:
            self.slots = torch.empty(self.world * self.slot, **f32)
            self.small = torch.empty(self.small_elems, **f32)
            self.my_slot = self.slots[self.rank * self.slot: (self.rank + 1) * self.slot]
        self.small.zero_()
        P_ = self.P
        self.views = {
            "dmeans3D": self.small[: 3 * P_].view(P_, 3),
            "dcov3D": self.small[3 * P_: 9 * P_].view(P_, 6),
            "dopacity": self.small[9 * P_: 10 * P_].view(P_, 1),
            "dcolors": self.my_slot[: 3 * P_].view(P_, 3),
        }
        shape = (P_, 3, self.K) if layout and layout.get("sh_channel_major") else (P_, self.K, 3)
        self.dsh = torch.empty(shape, **f32)

    def exchange_bytes(self) -> dict:
        """Bytes this rank sends over the interconnect per step (ring/NVLS lower bounds), for reports."""
        w = self.world
        return {"gather_recv": 4 * 3 * (self.P + 1) * (w - 1), "allreduce_payload": 4 * self.small_elems,
                "arena_allreduce_payload_replaced": 4 * self.P * (10 + 3 * self.K)}

    def kernels_per_step(self) -> int:
        """Kernels of this library the exchange adds to a step (the SH-gradient merge, + the NVLS all-reduce and,
        with barrier="nvls", two barrier kernels on the p2p transport); torch / NCCL kernels are not counted."""
        if self.world == 1 or self.transport != "p2p":
            return 1
        if self.signalled:
            return 2  # merge + all-reduce, both with their waits inside
        n = 1 + (1 if self.handles[1].multicast_ptr else 0)
        return n + (2 if self.barrier_kind == "nvls" else 0)

    def _barrier(self, channel: int) -> None:
        if self.barrier_kind == "nvls":
            from . import _cabi
            import ctypes as C

            self.epoch += 1
            stream = torch.cuda.current_stream(self.device)
            _cabi.check(_cabi.lib().ggrt_raster_nvls_barrier(
                C.c_void_p(self.bar_handle.multicast_ptr), C.c_void_p(self.bar.data_ptr()),
                (self.world * self.epoch) & 0xFFFFFFFF, C.c_void_p(stream.cuda_stream)), "nvls_barrier")
        else:
            self.handles[0].barrier(channel=channel)

    def _mark(self, name: str) -> None:
        """Phase boundary for `profile = True` (diagnostics): CUDA events on the current stream."""
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.device))
            self.marks.append((name, ev))

    def phase_ms(self) -> dict:
        """Durations between the marks of the last profiled run() (waits for the last event)."""
        if not self.marks:
            return {}
        self.marks[-1][1].synchronize()
        return {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self.marks[:-1], self.marks[1:])}

    def run(self, state: dict, grad_color: torch.Tensor, **backward_kw) -> dict:
        P_ = self.P
        self.marks = []
        self._mark("start")
        call = state["call"]
        if self.sinks is not None:  # p2p: the kernel pushes colour gradients + campos into every GPU's slot table
            out = self.backward_fn(state, grad_color, out={k: v for k, v in self.views.items() if k != "dcolors"},
                                   color_sinks=self.sinks, **backward_kw)
        else:
            out = self.backward_fn(state, grad_color, out=dict(self.views), compact=True, **backward_kw)
            self.my_slot[3 * P_: 3 * P_ + 3].copy_(call.campos.reshape(3))
        self._mark("backward_compact")
        if self.signalled:
            from . import _cabi
            import ctypes as C

            L = _cabi.lib()
            hs, hm = self.handles
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            side = self._side
            ctrl, bar, bar_mc = self.ctrl.data_ptr(), self.bar.data_ptr(), int(self.bar_handle.multicast_ptr)
            # the NVLink-bound sum of the small arena runs beside the HBM-bound merge; both kernels wait by themselves
            # for the arrival counter the ranks' backward kernels bump, the all-reduce also for its own completion
            side.wait_stream(main)
            _cabi.check(L.ggrt_raster_nvls_allreduce_signalled(
                C.c_void_p(hm.multicast_ptr), self.small_elems, self.rank, self.world, C.c_void_p(ctrl),
                C.c_void_p(bar), C.c_void_p(bar_mc + 4), C.c_void_p(bar + 4), C.c_void_p(ctrl + 8),
                C.c_void_p(side.cuda_stream)), "nvls_allreduce_signalled")
            clay = None
            if self.layout:
                clay = C.byref(_cabi.InputLayout(float(self.layout.get("scene_scale", 1.0)), 0,
                                                 int(bool(self.layout.get("sh_channel_major", False)))))
            _cabi.check(L.ggrt_raster_sh_gradient_merge_signalled(
                P_, self.deg, clay, C.c_void_p(call.means3D.data_ptr()), self.world, C.c_void_p(self.slots.data_ptr()),
                self.slot, self.world * self.slot, C.c_void_p(ctrl), C.c_void_p(bar), C.c_void_p(self.dsh.data_ptr()),
                C.c_void_p(main.cuda_stream)), "sh_gradient_merge_signalled")
            dsh = self.dsh
            self._mark("merge")
            main.wait_stream(side)
            self._mark("join_allreduce")
        elif self.world > 1 and self.transport == "p2p":
            from . import _cabi
            import ctypes as C

            hs, hm = self.handles
            self._barrier(0)  # every rank's colour gradients and small arena are complete
            self._mark("barrier_in")
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            side = self._side
            # the NVLink-bound sum of the small arena runs beside the HBM-bound merge
            side.wait_stream(main)
            if hm.multicast_ptr:  # NVLS available: the NVSwitch does the sum
                _cabi.check(_cabi.lib().ggrt_raster_nvls_allreduce_f32(
                    C.c_void_p(hm.multicast_ptr), self.small_elems, self.rank, self.world,
                    C.c_void_p(side.cuda_stream)), "nvls_allreduce")
            else:
                with torch.cuda.stream(side):
                    dist.all_reduce(self.small, op=dist.ReduceOp.SUM, group=self.group)
            slots = self.slots.view(self.world, self.slot)  # filled by the peers' kernels, local reads only
            dsh = self.merge_fn(call.means3D, self.deg, [slots[r, : 3 * P_] for r in range(self.world)],
                                [slots[r, 3 * P_: 3 * P_ + 3] for r in range(self.world)], out=self.dsh,
                                layout=self.layout)
            self._mark("merge")
            main.wait_stream(side)
            self._mark("join_allreduce")
            self._barrier(1)  # peers have finished reading this rank's buffers; the reduced arena is visible
            self._mark("barrier_out")
        else:
            if self.world > 1:
                dist.all_gather_into_tensor(self.slots, self.my_slot, group=self.group)
                self._mark("all_gather")
                dist.all_reduce(self.small, op=dist.ReduceOp.SUM, group=self.group)
                self._mark("all_reduce")
            slots = self.slots.view(self.world, self.slot)
            dsh = self.merge_fn(call.means3D, self.deg, [slots[r, : 3 * P_].view(P_, 3) for r in range(self.world)],
                                [slots[r, 3 * P_: 3 * P_ + 3] for r in range(self.world)], out=self.dsh,
                                layout=self.layout)
            self._mark("merge")
        return {"dmeans3D": self.views["dmeans3D"], "dcov3D": self.views["dcov3D"],
                "dopacity": self.views["dopacity"], "dsh": dsh, "dmeans2D": out["dmeans2D"]}


def make_exchange(P: int, sh_degree: int, device, group=None, prefer: str = "p2p", layout: Optional[dict] = None):
    """CompactGradientExchange with the preferred transport if EVERY rank can set it up (symmetric memory needs
    P2P access / a fabric between all GPUs), otherwise -- decided collectively, so no rank is left waiting in a
    rendezvous -- with the NCCL transport."""
    on = dist.is_available() and dist.is_initialized()
    if not on or dist.get_world_size(group) == 1 or prefer == "nccl":
        return CompactGradientExchange(P, sh_degree, device, group=group, transport="nccl", layout=layout)
    ex, ok = None, 1
    try:
        ex = CompactGradientExchange(P, sh_degree, device, group=group, transport=prefer, layout=layout)
    except Exception:  # noqa: BLE001 - any failure means "fall back", the reason is not actionable here
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        return ex
    del ex
    return CompactGradientExchange(P, sh_degree, device, group=group, transport="nccl", layout=layout)


class _ViewParallelRasterize(torch.autograd.Function):
    """This rank's view through the rasterizer; the backward sums the Gaussian gradients over all ranks' views
    (compact exchange) before handing them to autograd."""

    @staticmethod
    def forward(ctx, means3D, shs, opacities, cov3D_precomp, raster_settings, exchange, check):
        from . import rasterizer as R

        st = R.forward_raw(means3D, shs, None, opacities, cov3D_precomp, raster_settings, check=check)
        ctx.state = {k: st[k] for k in ("call", "radii", "geom", "img", "binning", "N", "capacity", "workspace")}
        ctx.exchange, ctx.check = exchange, check
        ctx.opacity_shape = tuple(opacities.shape)
        ctx.mark_non_differentiable(st["radii"], st["depth"])
        ctx.set_materialize_grads(False)
        return st["color"], st["radii"], st["depth"]

    @staticmethod
    def backward(ctx, grad_color, _grad_radii=None, _grad_depth=None):
        from . import rasterizer as R

        c = ctx.state["call"]
        if grad_color is None:  # every rank must still take part in the exchange
            grad_color = torch.zeros((3, c.H, c.W), dtype=torch.float32, device=c.device)
        if ctx.check == "lazy":
            R.check_pending(block=True)
        g = ctx.exchange.run(ctx.state, grad_color)
        return g["dmeans3D"], g["dsh"], g["dopacity"].reshape(ctx.opacity_shape), g["dcov3D"], None, None, None


class ViewParallelRasterizer(torch.nn.Module):
    """`GaussianRasterizer` for view-sharded training: every rank holds the same Gaussians and renders
    ITS target view (`raster_settings` differ per rank); `backward()` returns the gradients summed over the views of
    all ranks, exchanged by `exchange` (a CompactGradientExchange, see make_exchange) inside the backward -- no
    separate all-reduce of the 340 B/Gaussian gradient set afterwards.  Inputs as GGRt passes them: `shs` [P,K,3] and
    `cov3D_precomp` [P,6].  The returned gradient tensors are the exchange's own buffers, valid until the next
    backward through the same exchange (use them / step the optimiser before that, as a training loop does)."""

    def __init__(self, raster_settings, exchange: CompactGradientExchange, check: str = "lazy"):
        super().__init__()
        self.raster_settings, self.exchange, self.check = raster_settings, exchange, check

    def forward(self, means3D, opacities, shs, cov3D_precomp):
        from . import rasterizer as R

        key = (means3D.device.index, int(means3D.shape[0]), int(self.raster_settings.image_height),
               int(self.raster_settings.image_width))
        check = self.check if key in R._capacity_cache else "sync"  # the first frame of a shape sizes the pair buffer
        return _ViewParallelRasterize.apply(means3D, shs, opacities, cov3D_precomp, self.raster_settings,
                                            self.exchange, check)
