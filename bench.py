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
        "geometry+color": P * (40 + 12 * K) + 52 * P,
        "binning": 12 * N + 24 * N + 8 * T,  # emit + one read/one write of the pairs + ranges
        "render_forward": 4 * N + 36 * P + 8 * T + 20 * HW,
        "render_backward": 4 * N + 36 * P + 20 * HW + 44 * P,
        "preprocess_backward": P * (44 + 40 + 12 * K + 4) + P * (36 + 12 * K),
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


def run_reference(args, rank, world):
    """CPU arm: the oracle port (the reference's rasterizer is CUDA-only and not vendored), all host threads."""
    if rank != 0:
        return
    from oracle import c_oracle as co

    ri, g = make_inputs(args.workload, 0)
    cam = co.Camera(W=ri.image_width, H=ri.image_height, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, view=ri.viewmatrix,
                    proj=ri.projmatrix, campos=ri.campos, bg=ri.bg, deg=ri.sh_degree)

    def step():
        f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
        co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, g, sh=ri.shs)
        return f["bin"]["N"]

    for _ in range(args.warmup):
        N = step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        N = step()
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    P, H, W, desc = WORKLOADS[args.workload]
    sample = f"{args.steps} full frames (fwd+bwd) of the same workload, one frame per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "P": P, "H": H, "W": W, "N_pairs": int(N), "sh_degree": SH_DEGREE,
                   "note": "CPU oracle port (oracle/raster_oracle.c, OpenMP): the reference rasterizer is an "
                           "un-vendored CUDA-only package, no reference binary exists on this box"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": co.num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from ggrt_official_b200 import GaussianRasterizationSettings, GaussianRasterizer, _cabi
    from ggrt_official_b200 import rasterizer as R

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

    state = {}

    from ggrt_official_b200.view_parallel import GradientArena, make_exchange

    # Sum of the per-view Gaussian gradients over the GPUs, inside the timed step:
    #   arena   - ONE NCCL all-reduce of the contiguous [P, 3+6+1+3K] gradient arena (340 B / Gaussian)
    #   compact - exchange the [P,3] colour gradients + all-reduce [P,10], rebuild dL/dsh locally (sh_merge.cu)
    #   p2p     - compact over symmetric memory: in-kernel NVLink gather + NVLS multimem reduction, no NCCL
    #   auto    - p2p if every rank can set up symmetric memory, else compact (default)
    arena = GradientArena.allocate(P, K, dev) if world > 1 and args.exchange == "arena" else None
    exch = None
    if world > 1 and args.exchange != "arena":
        exch = make_exchange(P, SH_DEGREE, dev, prefer="nccl" if args.exchange == "compact" else "p2p")
        if args.exchange == "p2p" and exch.transport != "p2p":
            raise SystemExit("--exchange p2p: symmetric memory could not be set up on every rank")

    def step():
        st = R.forward_raw(means, shs, None, opac, cov, rs)
        if exch is not None:
            grads = exch.run(st, grad_img, want_camera=args.pose_grads)
        else:
            grads = R.backward_raw(st, grad_img, out=arena.views if arena else None, want_camera=args.pose_grads)
            if arena:
                arena.all_reduce()
        state["N"], state["max_tile_pairs"] = st["N"], st["max_tile_pairs"]
        return grads

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()

    # ---- timed region: K steps, L2 flushed before each, CUDA events on the launching stream -------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for a, b in evs:
        flush_l2()
        a.record()
        step()
        b.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    fps = world * args.steps / (total_ms * 1e-3)

    # ---- per-kernel durations (same process, same inputs, CUDA events inside the library) ---------
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
    N = state["N"]
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
    try:
        prof = json.loads((ROOT / "profiles" / "r1_ncu_kernels.json").read_text())
        if args.gaussians == 0 and prof.get("workload", "").startswith(WORKLOADS[args.workload][3][:2]):
            members = [k for k, grp in STAGE_GROUP.items() if grp == top]
            vals = [v["dram_traffic_bytes"] for name, v in prof["kernels"].items()
                    if any(name.startswith(m + "_kernel") for m in members)]
            if vals:
                traffic, traffic_src = int(sum(vals)), "profiles/r1_ncu_kernels.json (dram__bytes_read.sum + dram__bytes_write.sum)"
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

    # ---- end to end: host buffers -> GaussianRasterizer (autograd) -> image back on the host -------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_means, h_cov, h_opac, h_shs = pin(ri.means3D), pin(ri.cov3D), pin(ri.opacities), pin(ri.shs)
    h_img = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
    h_loss = torch.empty((), dtype=torch.float32).pin_memory()
    h2d = sum(x.numel() * 4 for x in (h_means, h_cov, h_opac, h_shs))
    d2h = h_img.numel() * 4 + 4
    rasterizer = GaussianRasterizer(rs)

    # Two sets of device staging tensors, allocated once: the H2D copy of step i+1's inputs is issued on a copy
    # stream while step i computes (the usual input prefetch of a training loop).  Every step still copies its
    # own 102 MB from pinned host memory and reads its image + loss back, all inside the timed region.
    d_sets = [[torch.empty_like(x, device=dev).requires_grad_() for x in (h_means, h_cov, h_opac, h_shs)]
              for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    filled = [torch.cuda.Event(), torch.cuda.Event()]   # set k holds fresh inputs
    free = [torch.cuda.Event(), torch.cuda.Event()]     # the step that used set k has finished with it
    main = torch.cuda.current_stream(dev)
    e2e_state = {"i": 0}

    def prefetch(k):
        copy_stream.wait_event(free[k])
        with torch.cuda.stream(copy_stream), torch.no_grad():
            for d_, h_ in zip(d_sets[k], (h_means, h_cov, h_opac, h_shs)):
                d_.copy_(h_, non_blocking=True)
        filled[k].record(copy_stream)

    for k in range(2):
        free[k].record(main)
    prefetch(0)

    def e2e_step():
        k = e2e_state["i"] & 1
        e2e_state["i"] += 1
        prefetch(k ^ 1)              # next step's inputs, overlapped with this step's kernels
        main.wait_event(filled[k])
        m, c, o, s = d_sets[k]
        for d_ in d_sets[k]:
            d_.grad = None
        m2 = torch.zeros_like(m, requires_grad=True)
        image, radii, _ = rasterizer(means3D=m, means2D=m2, shs=s, colors_precomp=None, opacities=o, cov3D_precomp=c)
        loss = (image * grad_img).sum()
        loss.backward()
        if world > 1:
            for p_ in (m, c, o, s):
                dist.all_reduce(p_.grad)
        free[k].record(main)
        h_img.copy_(image.detach(), non_blocking=True)
        h_loss.copy_(loss.detach(), non_blocking=True)

    for _ in range(5):
        e2e_step()
    sync_all()
    n_e2e = min(args.steps, 20)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_e2e)]
    for a, b in ev2:
        a.record()
        e2e_step()
        b.record()
    sync_all()
    e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)
    t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_fps = world * n_e2e / (float(t2.item()) * 1e-3)

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
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "P": P, "H": H, "W": W, "N_pairs": int(N),
                       "max_pairs_per_tile": int(state["max_tile_pairs"]), "sh_degree": SH_DEGREE,
                       "views_per_step": world, "pose_grads": bool(args.pose_grads), "l2": "flushed before every timed step (256 MiB write, then 256 MiB read so the step does not pay the flush's write-back)",
                       "parallelism": f"one target view per GPU x{world}" + (
                           "" if world == 1 else
                           ", NCCL all-reduce of the Gaussian gradient arena" if exch is None else
                           f", compact gradient exchange ({exch.transport}): gather [P,3] colour gradients + "
                           "all-reduce [P,10], dL/dsh rebuilt per GPU"),
                       "exchange_bytes": None if exch is None else exch.exchange_bytes()},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": n_e2e},
            # this library's kernels per step: 8 of the fwd+bwd path, + the SH-gradient merge (compact exchange), + the
            # NVLS all-reduce kernel (p2p transport); torch / NCCL kernels (barriers, collectives) are not counted
            "gpu_launches": (8 + (0 if exch is None else 2 if exch.transport == "p2p" else 1)) * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "gpu_baseline": gpu_base,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
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
