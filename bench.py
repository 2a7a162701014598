#!/usr/bin/env python
"""Benchmark of the rasterizer hot path: forward+backward frames/s on BASELINE.json's configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4] [--impl ours|reference]

N>1 is launched by the driver under torch.distributed.run (one rank per GPU, NCCL):
every rank renders a different target view of the same Gaussians (weak scaling: one frame
per GPU per step) and the Gaussian gradients are all-reduced inside the timed step.

One JSON line on stdout (rank 0).  `value` = frames/s with inputs resident in HBM, timed
through the C ABI; `e2e` = the same metric through GaussianRasterizer (autograd) with the
inputs copied from pinned host memory and the image read back every step; `roofline` =
algorithmic bytes of the dominant kernel / its CUDA-event duration; `cpu_baseline` = the
CPU oracle (a port: the reference has no CPU path) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (P, H, W, description) -- SURVEY.md 8d
    "c1": (10_000, 256, 256, "C1: 10K Gaussians, 256x256 (correctness config)"),
    "c2": (300_000, 756, 1008, "C2: LLFF fern shape, 4 source views -> 300K Gaussians, 1008x756, SH degree 4, fwd+bwd"),
    "c3": (921_600, 1280, 1920, "C3: Waymo scene-019 shape, 921.6K Gaussians, 1920x1280, SH degree 4, fwd+bwd"),
    "c4": (600_000, 756, 1008, "C4: 8 source views -> 600K Gaussians, 1008x756, SH degree 4, fwd+bwd"),
}
METRIC = "rasterizer fwd+bwd frames/sec @1008x756, 300K Gaussians"
SH_DEGREE = 4


def alg_bytes(P, N, H, W, K=25):
    """Algorithmic HBM bytes per launch of each kernel group (restated in DESIGN.md)."""
    T = ((W + 15) // 16) * ((H + 15) // 16)
    HW = H * W
    return {
        "geometry+color": P * (40 + 12 * K) + 52 * P + 36 * P,  # + the colour Jacobian kept for the backward
        "binning": 12 * N + 24 * N + 8 * T,  # emit + one read/one write of the pairs + ranges
        "render_forward": 4 * N + 36 * P + 8 * T + 20 * HW,
        "render_backward": 4 * N + 36 * P + 20 * HW + 44 * P,
        # round 2: the SH table (12K B) is no longer read here, the 36-byte colour Jacobian of the forward replaces it
        "preprocess_backward": P * (44 + 40 + 36 + 4) + P * (36 + 12 * K),
    }


STAGE_GROUP = {
    "geometry": "geometry+color", "scan_tiles": "binning", "color": "geometry+color", "emit": "binning",
    "sort_tiles": "binning", "render_forward": "render_forward", "render_backward": "render_backward",
    "preprocess_backward": "preprocess_backward",
}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region with NVML (a 2 ms polling thread: the timed
    region is only tens of milliseconds long, too short for `nvidia-smi -lms`)."""

    REASONS = {  # nvmlClocksEventReasons bit -> name used by B200_PROFILING.md
        0x0000000000000008: "hw_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000004: "sw_power_cap",
    }

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples, self.bits = [], 0
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.err = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = f"NVML unavailable: {e}"
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception as e:  # pragma: no cover
                self.err = str(e)
                return
            time.sleep(0.002)

    def stop(self) -> dict:
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [self.err or "not started"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        reasons = sorted(n for bit, n in self.REASONS.items() if self.bits & bit)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": reasons}


def make_inputs(workload: str, rank: int):
    from ggrt_official_b200.synthetic import SEED, image_gradient, make_scene, small_se3, to_raster_inputs

    P, H, W, _ = WORKLOADS[workload]
    scene = make_scene(P, H, W, sh_degree=SH_DEGREE, seed=SEED)
    if rank > 0:  # another target view of the same Gaussians (independent unit: cuda_splatting.py:93-127)
        rng = np.random.default_rng(SEED + 1000 + rank)
        scene.extrinsics = (scene.extrinsics.astype(np.float64) @ small_se3(rng).astype(np.float64)).astype(np.float32)
    return to_raster_inputs(scene), image_gradient(H, W, seed=SEED + rank)


def base_config(args, N_pairs, max_pairs):
    """The workload description both arms print (identical key set, so the driver can compare the lines)."""
    P, H, W, desc = WORKLOADS[args.workload]
    return {"workload": desc, "P": P, "H": H, "W": W, "N_pairs": int(N_pairs), "max_pairs_per_tile": int(max_pairs),
            "sh_degree": SH_DEGREE, "views_per_step": 1, "pose_grads": bool(args.pose_grads)}


def run_reference(args, rank, world):
    """CPU arm: the oracle port (the reference's rasterizer is CUDA-only and not vendored), all host threads.
    Under torchrun rank 0 alone runs it; torchrun exports OMP_NUM_THREADS=1, which is overridden here."""
    if rank != 0:
        return
    from oracle import c_oracle as co

    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    co.set_num_threads(ncpu)
    ri, g = make_inputs(args.workload, 0)
    cam = co.Camera(W=ri.image_width, H=ri.image_height, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, view=ri.viewmatrix,
                    proj=ri.projmatrix, campos=ri.campos, bg=ri.bg, deg=ri.sh_degree)

    def step():
        f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
        co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs, want_camera=args.pose_grads)
        cnt = f["bin"]["ranges"][:, 1] - f["bin"]["ranges"][:, 0]
        return f["bin"]["N"], int(cnt.max())

    for _ in range(args.warmup):
        N, mx = step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        N, mx = step()
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    sample = f"{args.steps} full frames (fwd+bwd) of the same workload, one frame per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args, N, mx),
        "details": {"note": "CPU oracle port (oracle/raster_oracle.c, OpenMP): the reference rasterizer is an "
                            "un-vendored CUDA-only package, no reference binary exists on this box",
                    "gpus_used": 0, "host_threads": co.num_threads()},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": co.num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_loop(fn, n, dev, before=None):
    """n calls of fn, each bracketed by CUDA events on the current stream (`before` runs untimed in front of each);
    returns the per-call milliseconds."""
    import torch

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        if before is not None:
            before()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    return [a.elapsed_time(b) for a, b in evs]


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from ggrt_official_b200 import GaussianRasterizationSettings, GaussianRasterizer, _cabi
    from ggrt_official_b200 import rasterizer as R
    from ggrt_official_b200.graph import CapturedStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P, H, W, desc = WORKLOADS[args.workload]
    ri, g_np = make_inputs(args.workload, rank)
    K = (SH_DEGREE + 1) ** 2

    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    means, cov, opac, shs = t(ri.means3D), t(ri.cov3D), t(ri.opacities), t(ri.shs)
    grad_img = t(g_np)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(ri.bg), scale_modifier=1.0,
        viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix), sh_degree=ri.sh_degree, campos=t(ri.campos),
        prefiltered=False)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    flush_src = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MiB, read-only
    flush_sink = torch.zeros((), dtype=torch.float32, device=dev)

    def flush_l2():
        """Evict the working set: write 256 MiB, then read 256 MiB so the dirty lines of the write are
        themselves written back before the timed step (otherwise the step pays for the flush's write-back)."""
        flush_buf.zero_()
        flush_sink.copy_(flush_src.sum())

    from ggrt_official_b200.view_parallel import GradientArena, ViewParallelRasterizer, make_exchange

    # Sum of the per-view Gaussian gradients over the GPUs, inside the timed step:
    #   arena   - ONE NCCL all-reduce of the contiguous [P, 3+6+1+3K] gradient arena (340 B / Gaussian)
    #   compact - exchange the [P,3] colour gradients + all-reduce [P,10], rebuild dL/dsh locally (sh_merge.cu)
    #   p2p     - compact over symmetric memory: in-kernel NVLink push + NVLS multimem reduction, no NCCL
    #   auto    - p2p if every rank can set up symmetric memory, else compact (default)
    arena = GradientArena.allocate(P, K, dev) if world > 1 and args.exchange == "arena" else None
    exch = None
    if world > 1 and args.exchange != "arena":
        exch = make_exchange(P, SH_DEGREE, dev, prefer="nccl" if args.exchange == "compact" else "p2p")
        if args.exchange == "p2p" and exch.transport != "p2p":
            raise SystemExit("--exchange p2p: symmetric memory could not be set up on every rank")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- the step: forward + backward (+ exchange).  Default: captured once into a CUDA graph (graph.CapturedStep),
    # one launch per step and no host wait; --launch eager issues the same kernels call by call with the pair-buffer
    # check deferred (check="lazy"), so the host does not wait for the device there either -------------------------
    st0 = R.forward_raw(means, shs, None, opac, cov, rs)  # exact first frame: sizes the pair buffer estimate
    N, max_pairs = st0["N"], st0["max_tile_pairs"]
    del st0
    launch, cap, cap_err = args.launch, None, None
    if launch == "auto":
        launch = "graph"
    if launch == "graph" and arena is None:
        try:
            cap = CapturedStep(means, shs, None, opac, cov, rs, grad_img, exchange=exch, want_camera=args.pose_grads)
        except Exception as e:  # noqa: BLE001 - fall back to eager launches, and say so in the line
            cap_err, cap = repr(e)[:300], None
            torch.cuda.synchronize()
    ok = torch.tensor([1 if (cap is not None or launch != "graph" or arena is not None) else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        cap = None
    if cap is None:
        launch = "eager"

    def step():
        if cap is not None:
            cap.replay()
            return
        st = R.forward_raw(means, shs, None, opac, cov, rs, check="lazy")
        if exch is not None:
            exch.run(st, grad_img, want_camera=args.pose_grads)
        else:
            R.backward_raw(st, grad_img, out=arena.views if arena else None, want_camera=args.pose_grads)
            if arena:
                arena.all_reduce()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()

    # ---- timed region: K steps, L2 flushed before each, CUDA events on the launching stream -------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync_all()
    w0 = time.perf_counter()
    per_step = time_loop(step, args.steps, dev, before=flush_l2)
    wall_ms = 1e3 * (time.perf_counter() - w0)
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    if cap is not None:
        N, max_pairs = cap.check()  # raises if a replay overflowed the captured pair buffer
    else:
        R.check_pending(block=True)
    total_ms = sum(per_step)
    rank_ms = [total_ms]
    rank_wall = [wall_ms]
    if world > 1:
        tt = torch.tensor([total_ms, wall_ms], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)
        rank_ms = [float(x[0]) for x in allt]
        rank_wall = [float(x[1]) for x in allt]
    total_ms = max(rank_ms)
    fps = world * args.steps / (total_ms * 1e-3)

    # ---- per-kernel durations (same process, same inputs, CUDA events inside the library, eager launches with the
    # colour kernel back on the main stream) ------------------------------------------------------------------
    stage_ms = {}
    _cabi.profile_enable(True)
    nprof = min(args.steps, 20)
    for _ in range(nprof):
        flush_l2()
        st = R.forward_raw(means, shs, None, opac, cov, rs)
        fw = _cabi.profile_read()
        R.backward_raw(st, grad_img)
        bw = _cabi.profile_read()
        for k in fw:
            v = fw[k] if k not in ("render_backward", "preprocess_backward") else bw[k]
            stage_ms[k] = stage_ms.get(k, 0.0) + v / nprof
    _cabi.profile_enable(False)
    ab = alg_bytes(P, N, H, W, K)
    group_ms = {}
    for k, v in stage_ms.items():
        group_ms[STAGE_GROUP[k]] = group_ms.get(STAGE_GROUP[k], 0.0) + v
    top = max(group_ms, key=group_ms.get)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # measured DRAM traffic of the same kernel(s) from the committed ncu --set full capture (per launch)
    traffic, traffic_src = None, None
    for prof_name in ("r2_ncu_kernels.json", "r1_ncu_kernels.json"):
        try:
            prof = json.loads((ROOT / "profiles" / prof_name).read_text())
            if args.gaussians == 0 and prof.get("workload", "").startswith(WORKLOADS[args.workload][3][:2]):
                members = [k for k, grp in STAGE_GROUP.items() if grp == top]
                vals = [v["dram_traffic_bytes"] for name, v in prof["kernels"].items()
                        if any(name.startswith(m + "_kernel") for m in members)]
                if vals:
                    traffic = int(sum(vals))
                    traffic_src = f"profiles/{prof_name} (dram__bytes_read.sum + dram__bytes_write.sum)"
                    break
        except Exception:
            pass
    achieved = ab[top] / (group_ms[top] * 1e-3) / 1e9
    path_bytes = sum(ab.values())
    kernel_sum_ms = sum(stage_ms.values())
    roofline = {
        "bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": group_ms[top], "algorithmic_bytes": ab[top],
        "kernel_share_of_step": group_ms[top] / kernel_sum_ms if kernel_sum_ms else None,
        "stage_ms": {k: round(v, 5) for k, v in stage_ms.items()},
        "whole_path": {"algorithmic_bytes": path_bytes, "achieved": path_bytes / (total_ms / args.steps * 1e-3) / 1e9,
                       "frac": path_bytes / (total_ms / args.steps * 1e-3) / 1e9 / peak},
    }

    # ---- end to end: pinned host buffers -> public API (autograd) -> image + loss back on the host ---------------
    # N = 1: GaussianRasterizer, every step copies its 102 MB of Gaussians from pinned memory.
    # N > 1: the Gaussians are the same for every rank, so rank r uploads rows [r P/N, (r+1) P/N) over ITS PCIe link
    # and the ranks all-gather the shards over NVLink (in place, on the copy stream, overlapped with the previous
    # step); the step itself is ViewParallelRasterizer: forward, backward and the compact gradient exchange.
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host_full = (ri.means3D, ri.cov3D, ri.opacities, ri.shs)
    chunk = (P + world - 1) // world
    lo, hi = rank * chunk, min(P, (rank + 1) * chunk)
    h_in = [pin(a[lo:hi]) for a in host_full]
    h_img = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
    h_loss = torch.empty((), dtype=torch.float32).pin_memory()
    h2d = sum(x.numel() * 4 for x in h_in)
    d2h = h_img.numel() * 4 + 4
    if world > 1:
        rasterizer = ViewParallelRasterizer(rs, exch) if exch is not None else GaussianRasterizer(rs)
    else:
        rasterizer = GaussianRasterizer(rs)

    # Two sets of device staging tensors, allocated once (padded to world * chunk rows so that the all-gather has
    # equal shards): the upload of step i+1 is issued on a copy stream while step i computes (the usual input
    # prefetch of a training loop).  Every step still moves its own bytes and reads its image + loss back.
    d_sets = [[torch.empty((world * chunk,) + tuple(a.shape[1:]), dtype=torch.float32, device=dev) for a in host_full]
              for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    filled = [torch.cuda.Event(), torch.cuda.Event()]   # set k holds fresh inputs
    free = [torch.cuda.Event(), torch.cuda.Event()]     # the step that used set k has finished with it
    main = torch.cuda.current_stream(dev)
    e2e_state = {"i": 0}

    def prefetch(k):
        copy_stream.wait_event(free[k])
        with torch.cuda.stream(copy_stream), torch.no_grad():
            for d_, h_ in zip(d_sets[k], h_in):
                d_[lo:hi].copy_(h_, non_blocking=True)
            if world > 1:
                for d_ in d_sets[k]:
                    dist.all_gather_into_tensor(d_, d_[rank * chunk:(rank + 1) * chunk])
        filled[k].record(copy_stream)

    for k in range(2):
        free[k].record(main)
    prefetch(0)

    def e2e_step():
        k = e2e_state["i"] & 1
        e2e_state["i"] += 1
        prefetch(k ^ 1)              # next step's inputs, overlapped with this step's kernels
        main.wait_event(filled[k])
        m, c, o, s = [d_[:P].detach().requires_grad_() for d_ in d_sets[k]]
        if isinstance(rasterizer, ViewParallelRasterizer):
            image, radii, _ = rasterizer(means3D=m, opacities=o, shs=s, cov3D_precomp=c)
        else:
            m2 = torch.zeros_like(m, requires_grad=True)
            image, radii, _ = rasterizer(means3D=m, means2D=m2, shs=s, colors_precomp=None, opacities=o,
                                         cov3D_precomp=c)
        loss = (image * grad_img).sum()
        loss.backward()
        if world > 1 and exch is None:
            for p_ in (m, c, o, s):
                dist.all_reduce(p_.grad)
        free[k].record(main)
        h_img.copy_(image.detach(), non_blocking=True)
        h_loss.copy_(loss.detach(), non_blocking=True)

    for _ in range(5):
        e2e_step()
    sync_all()
    n_e2e = min(args.steps, 20)
    e2e_ms = sum(time_loop(e2e_step, n_e2e, dev))
    sync_all()
    t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_fps = world * n_e2e / (float(t2.item()) * 1e-3)

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extras = measure_extras(args, ri, g_np, dev, flush_l2)

    # ---- CPU baseline: the oracle port on a bounded sample (rank 0, N=1 only) ----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle as co

        cam = co.Camera(W=W, H=H, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, view=ri.viewmatrix, proj=ri.projmatrix,
                        campos=ri.campos, bg=ri.bg, deg=ri.sh_degree)

        def cpu_frame():
            f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
            co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g_np, sh=ri.shs)

        t0 = time.perf_counter()
        cpu_frame()
        one = time.perf_counter() - t0
        n = int(min(50, max(2, round(12.0 / max(one, 1e-3)))))
        t0 = time.perf_counter()
        for _ in range(n):
            cpu_frame()
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "frames/s", "cores": co.num_threads(), "kind": "port",
               "sample": f"{n} full frames (fwd+bwd) of the same workload after one warm-up frame"}

    # ---- GPU baseline: the upstream kernel STRUCTURE restated for sm_100a (baseline/upstream_structure.cu; NOT the
    # reference binary, whose source is unavailable offline), timed on this GPU by tools/bench_upstream.py in a
    # subprocess so that a failure there cannot disturb this process (rank 0, N=1 only) ----------------------------
    gpu_base = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline and args.gaussians == 0:
        try:
            res = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_upstream.py"), "--workload", args.workload,
                                  "--steps", "20"], capture_output=True, text=True, timeout=240)
            rows = [l for l in res.stdout.splitlines() if l.startswith("{")]
            if res.returncode == 0 and rows:
                d = json.loads(rows[-1])
                gpu_base = {"kind": "upstream-structure restatement on this GPU (not the reference binary)",
                            "value": d["upstream_structure_fps"], "unit": "frames/s", "ms_per_step": d["upstream_structure_ms"],
                            "product_ms_same_harness": d["product_ms"], "speedup": d["speedup"],
                            "same_results": {"radii_equal": d["radii_equal"], "color_max_abs_diff": d["color_max_abs_diff"],
                                             "grad_max_rel_diff": max(d["grad_max_rel_diff"].values())}}
            else:
                gpu_base = {"error": (res.stderr or res.stdout)[-300:]}
        except Exception as e:  # noqa: BLE001 - a measurement aid must never fail the bench
            gpu_base = {"error": repr(e)[:300]}

    if rank == 0:
        cfg = base_config(args, N, max_pairs)
        cfg["views_per_step"] = world
        n_exch = 0 if exch is None else exch.kernels_per_step()
        line = {
            "impl": "ours", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "details": {
                "launch": ("one CUDA graph replay per step (graph.CapturedStep: forward + backward"
                           + (" + gradient exchange" if exch is not None else "") + ")") if cap is not None else
                          "eager kernel launches, pair-buffer check deferred (check='lazy')",
                "graph_capture_error": cap_err,
                "l2": "flushed before every timed step (256 MiB write, then 256 MiB read so the step does not pay "
                      "the flush's write-back)",
                "parallelism": f"one target view per GPU x{world}" + (
                    "" if world == 1 else
                    ", NCCL all-reduce of the Gaussian gradient arena" if exch is None else
                    f", compact gradient exchange ({exch.transport}): push [P,3] colour gradients + "
                    "all-reduce [P,10], dL/dsh rebuilt per GPU"),
                "exchange_bytes": None if exch is None else exch.exchange_bytes(),
                # skew between the ranks: device time of the timed steps vs the host's wall clock around them
                "rank_gpu_ms_per_step": [round(x / args.steps, 5) for x in rank_ms],
                "rank_wall_ms_per_step_incl_l2_flush": [round(x / args.steps, 5) for x in rank_wall],
                "e2e_api": type(rasterizer).__name__ + (
                    "" if world == 1 else ": per-rank upload of a 1/N shard + NVLink all-gather (NCCL) of the inputs"),
            },
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": n_e2e},
            # this library's kernels per step: 8 of the fwd+bwd path, the dL/dsh writer beside the per-Gaussian backward
            # when one GPU writes the full SH gradient itself (with an exchange the merge kernel writes it), + the
            # exchange's; torch / NCCL kernels (collectives, L2 flush) are not counted
            "gpu_launches": (8 + (1 if exch is None else 0) + n_exch) * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "gpu_baseline": gpu_base,
            "clocks": clocks,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_extras(args, ri, g_np, dev, flush_l2):
    """The two extra lines SURVEY.md 8d asks for, at N=1: the depth pass (degree-0 rasterization, fwd+bwd through the
    C ABI) and the path THROUGH THE CALLER -- the decoder-level mirror of the unmodified reference glue
    (render_cuda + render_depth_cuda incl. their host syncs and copies, cuda_splatting.py:49-128, :227-269), plus
    this package's fused and copy-free variants of it.  Frames/s, fwd+bwd, CUDA events."""
    import torch

    from ggrt_official_b200 import rasterizer as R
    from ggrt_official_b200.decoder import DecoderSplattingCUDA, Gaussians
    from ggrt_official_b200.synthetic import SEED, make_scene

    P, H, W, _ = WORKLOADS[args.workload]
    out = {}
    n = 20
    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    try:
        # depth pass: what render_depth_cuda rasterizes -- SH degree 0 "colours" (K = 1) of the same geometry
        from ggrt_official_b200 import GaussianRasterizationSettings

        means, cov, opac = t(ri.means3D), t(ri.cov3D), t(ri.opacities)
        sh0 = t(np.ascontiguousarray(np.repeat(ri.shs[:, :1, :1], 3, axis=2)))
        rs0 = GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(np.zeros(3, np.float32)),
            scale_modifier=1.0, viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix), sh_degree=0,
            campos=t(ri.campos), prefiltered=False)
        g = t(g_np)

        def depth_step():
            st = R.forward_raw(means, sh0, None, opac, cov, rs0, check="lazy")
            R.backward_raw(st, g)

        R.forward_raw(means, sh0, None, opac, cov, rs0)
        for _ in range(3):
            depth_step()
        ms = time_loop(depth_step, n, dev, before=flush_l2)
        R.check_pending(block=True)
        out["depth_pass"] = {"value": 1e3 * n / sum(ms), "unit": "frames/s", "ms_per_step": sum(ms) / n,
                             "what": "SH degree 0 (K=1) rasterization of the same Gaussians, fwd+bwd, C ABI, L2 flushed"}
    except Exception as e:  # noqa: BLE001
        out["depth_pass"] = {"error": repr(e)[:300]}
    try:
        # crop training (finetune_ggrt_stable.py:126-142): dL/dimage is zero outside one crop -- here one quadrant -- so the
        # backward render kernel skips three quarters of the tiles; same Gaussians, same forward
        from ggrt_official_b200.synthetic import image_gradient

        shs = t(ri.shs)
        rs4 = rs0._replace(sh_degree=ri.sh_degree)
        gq = t(image_gradient(H, W, seed=SEED, quadrant_only=True))

        def crop_step():
            st = R.forward_raw(means, shs, None, opac, cov, rs4, check="lazy", prezero_scratch=True)
            R.backward_raw(st, gq)

        R.forward_raw(means, shs, None, opac, cov, rs4)
        for _ in range(3):
            crop_step()
        ms = time_loop(crop_step, n, dev, before=flush_l2)
        R.check_pending(block=True)
        out["crop_gradient"] = {"value": 1e3 * n / sum(ms), "unit": "frames/s", "ms_per_step": sum(ms) / n,
                                "what": "the metric's workload with dL/dimage non-zero in one quadrant only (crop training), "
                                        "fwd+bwd, C ABI, eager launches, L2 flushed"}
    except Exception as e:  # noqa: BLE001
        out["crop_gradient"] = {"error": repr(e)[:300]}
    try:
        sc = make_scene(P, H, W, sh_degree=SH_DEGREE, seed=SEED)
        leaves = dict(means=t(sc.means)[None].requires_grad_(), covariances=t(sc.covariances)[None].requires_grad_(),
                      harmonics=t(sc.harmonics)[None].requires_grad_(), opacities=t(sc.opacities)[None].requires_grad_())
        extr, intr = t(sc.extrinsics)[None, None], t(sc.intrinsics)[None, None]
        near, far = torch.full((1, 1), sc.near, device=dev), torch.full((1, 1), sc.far, device=dev)
        wc = t(g_np)[None, None]
        wd = torch.randn(1, 1, H, W, device=dev) / (H * W)
        res = {}
        for name, fused, fast, mode in (("color_only", False, False, None), ("color_and_depth_two_pass", False, False, "depth"),
                                        ("color_and_depth_fused", True, False, "depth"),
                                        ("color_and_depth_fast_glue", True, True, "depth"),
                                        ("color_and_depth_device_glue", True, "device", "depth")):
            dec = DecoderSplattingCUDA(fused_depth=fused, fast_glue=bool(fast), device_glue=(fast == "device"))

            def dstep():
                for v in leaves.values():
                    v.grad = None
                r = dec(Gaussians(**leaves), extr, intr, near, far, (H, W), depth_mode=mode)
                loss = (r.color * wc).sum()
                if mode is not None:
                    loss = loss + (r.depth * wd).sum()
                loss.backward()

            for _ in range(3):
                dstep()
            ms = time_loop(dstep, n, dev)
            res[name] = {"value": 1e3 * n / sum(ms), "unit": "frames/s", "ms_per_step": sum(ms) / n}
        # several target views of the same Gaussians in one decoder call (GGRt renders 1-4 target views per step): the
        # sync-free glue with the views queued on one stream, on two streams, and with the pair-buffer check deferred
        try:
            nv = 4
            rng = np.random.default_rng(SEED + 5)
            from ggrt_official_b200.synthetic import small_se3

            E4 = np.stack([sc.extrinsics.astype(np.float64) @ small_se3(rng).astype(np.float64) for _ in range(nv)]).astype(np.float32)
            extr4, intr4 = t(E4)[None], t(np.stack([sc.intrinsics] * nv))[None]
            near4, far4 = torch.full((1, nv), sc.near, device=dev), torch.full((1, nv), sc.far, device=dev)
            wc4 = t(g_np)[None, None].expand(1, nv, 3, H, W)
            wd4 = wd.expand(1, nv, H, W)
            multi = {}
            for name, streams, check in (("one_stream", 1, "sync"), ("two_streams", 2, "sync"),
                                         ("one_stream_lazy_check", 1, "lazy"), ("two_streams_lazy_check", 2, "lazy")):
                dec = DecoderSplattingCUDA(device_glue=True, view_streams=streams)
                R.AUTOGRAD_CHECK = check

                def mstep():
                    for v in leaves.values():
                        v.grad = None
                    r = dec(Gaussians(**leaves), extr4, intr4, near4, far4, (H, W), depth_mode="depth")
                    ((r.color * wc4).sum() + (r.depth * wd4).sum()).backward()

                for _ in range(3):
                    mstep()
                ms = time_loop(mstep, 10, dev)
                R.check_pending(block=True)
                multi[name] = {"value": 1e3 * 10 * nv / sum(ms), "unit": "frames/s", "ms_per_call": sum(ms) / 10}
            R.AUTOGRAD_CHECK = "sync"
            # the same call as ONE captured CUDA-graph launch (graph.CapturedViews: compact per-view colour gradients,
            # one SH-gradient merge for all views, views on two streams inside the graph; fixed loss gradients)
            try:
                from ggrt_official_b200.graph import CapturedViews

                for name, streams in (("captured_graph_one_stream", 1), ("captured_graph_two_streams", 2)):
                    cap = CapturedViews(extr4[0], intr4[0], near4[0], far4[0], (H, W), torch.zeros(nv, 3, device=dev),
                                        leaves["means"][0], leaves["covariances"][0], leaves["harmonics"][0],
                                        leaves["opacities"][0], grad_color=wc4[0].contiguous(),
                                        grad_depth=wd4[0].contiguous(), streams=streams)
                    for _ in range(3):
                        cap.replay()
                    ms = time_loop(cap.replay, 10, dev)
                    cap.check()
                    multi[name] = {"value": 1e3 * 10 * nv / sum(ms), "unit": "frames/s", "ms_per_call": sum(ms) / 10}
                    del cap
            except Exception as e:  # noqa: BLE001
                multi["captured_graph"] = {"error": repr(e)[:300]}
            multi["what"] = f"{nv} target views of the same Gaussians per decoder call (colour + depth, fwd+bwd through autograd)"
            res["multi_view_device_glue"] = multi
        except Exception as e:  # noqa: BLE001
            R.AUTOGRAD_CHECK = "sync"
            res["multi_view_device_glue"] = {"error": repr(e)[:300]}
        res["what"] = ("DecoderSplattingCUDA (mirror of decoder_splatting_cuda.py:29-85 driving render_cuda / "
                       "render_depth_cuda, same host syncs and copies as the reference glue), 1 view, fwd+bwd through "
                       "autograd, inputs resident, no L2 flush")
        out["through_caller"] = res
    except Exception as e:  # noqa: BLE001
        out["through_caller"] = {"error": repr(e)[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true",
                    help="skip the upstream-structure GPU baseline (tools/bench_upstream.py, run in a subprocess)")
    ap.add_argument("--gaussians", type=int, default=0,
                    help="override the Gaussian count of the workload (BASELINE config 5: 50K..2M sweep at 1008x756)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "arena", "compact", "p2p"],
                    help="multi-GPU gradient exchange (N>1 only), see run_ours")
    ap.add_argument("--launch", default="auto", choices=["auto", "graph", "eager"],
                    help="how the timed step is issued: one CUDA graph replay, or eager kernel launches with the "
                         "pair-buffer check deferred; auto (default) = graph (measured on 8xB200, DESIGN.md section 6: "
                         "0.355 vs 0.367 ms at N=1, 0.433 vs 0.453 ms at N=8)")
    ap.add_argument("--no-extras", action="store_true", help="skip the depth-pass / through-the-caller lines (N=1)")
    ap.add_argument("--pose-grads", action="store_true",
                    help="also compute dL/d(viewmatrix, projmatrix, campos) in the backward (BASELINE config 3)")
    args = ap.parse_args()
    if args.gaussians > 0:
        P0, H0, W0, d0 = WORKLOADS[args.workload]
        WORKLOADS[args.workload] = (args.gaussians, H0, W0, f"C5 sweep point: {args.gaussians} Gaussians, {W0}x{H0}, SH degree 4, fwd+bwd")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
