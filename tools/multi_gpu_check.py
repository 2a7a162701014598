"""Multi-GPU check of the view-sharded gradient exchange (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/multi_gpu_check.py [--workload c2] [--iters 20]

Every rank renders another target view of the same Gaussians.  The reference result is the NCCL all-reduce of the
full gradient arena; the compact exchange (NCCL transport, then symmetric-memory P2P + NVLS transport) must
reproduce it within float32 summation noise.  Prints one JSON line per variant (rank 0) with the step time
(fwd + bwd + exchange, CUDA events, max over ranks, L2 flushed before each step)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import traceback
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from ggrt_official_b200 import GaussianRasterizationSettings  # noqa: E402
from ggrt_official_b200 import rasterizer as R  # noqa: E402
from ggrt_official_b200.view_parallel import CompactGradientExchange, GradientArena  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--check", action="store_true", help="exit non-zero unless every variant ran and matched the arena sum")
    ap.add_argument("--gaussians", type=int, default=0, help="override the Gaussian count of the workload")
    args = ap.parse_args()
    if args.gaussians:
        P0, H0, W0, d0 = bench.WORKLOADS[args.workload]
        bench.WORKLOADS[args.workload] = (args.gaussians, H0, W0, d0)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P, H, W, desc = bench.WORKLOADS[args.workload]
    ri, g_np = bench.make_inputs(args.workload, rank)
    K = (bench.SH_DEGREE + 1) ** 2
    t = lambda a: torch.tensor(np.asarray(a), device=dev)
    means, cov, opac, shs, grad_img = t(ri.means3D), t(ri.cov3D), t(ri.opacities), t(ri.shs), t(g_np)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(ri.bg), scale_modifier=1.0,
        viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix), sh_degree=ri.sh_degree, campos=t(ri.campos),
        prefiltered=False)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def say(obj):
        if rank == 0:
            print(json.dumps(obj), flush=True)

    arena = GradientArena.allocate(P, K, dev)

    def step_arena():
        st = R.forward_raw(means, shs, None, opac, cov, rs)
        R.backward_raw(st, grad_img, out=arena.views)
        arena.all_reduce()
        return arena.views

    def timed(step):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        tot = 0.0
        for _ in range(args.iters):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        tt = torch.tensor([tot / args.iters], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    ref = {k: v.clone() for k, v in step_arena().items()}
    torch.cuda.synchronize()
    say({"variant": "arena_allreduce", "world": world, "workload": args.workload, "ms_per_step": timed(step_arena),
         "exchange_bytes": int(arena.flat.numel() * 4)})

    def step_local():  # no exchange at all: the compute floor
        st = R.forward_raw(means, shs, None, opac, cov, rs)
        R.backward_raw(st, grad_img, out=arena.views)

    say({"variant": "no_exchange", "world": world, "ms_per_step": timed(step_local)})

    failures = []
    for transport in ("nccl", "p2p+signal", "p2p+signal+graph", "p2p+torch", "p2p+nvls"):
        try:
            parts = transport.split("+")
            ex = CompactGradientExchange(P, bench.SH_DEGREE, dev, transport=parts[0],
                                         barrier=parts[1] if len(parts) > 1 else "torch")

            def step_compact():
                st = R.forward_raw(means, shs, None, opac, cov, rs)
                return ex.run(st, grad_img)

            if "graph" in transport:  # the whole step (forward + backward + exchange) as ONE CUDA graph launch
                from ggrt_official_b200.graph import CapturedStep

                cap = CapturedStep(means, shs, None, opac, cov, rs, grad_img, exchange=ex)

                def step_compact():  # noqa: F811
                    cap.replay()
                    return cap.grads

            got = step_compact()
            torch.cuda.synchronize()
            errs = {}
            for k in ("dmeans3D", "dcov3D", "dopacity", "dsh"):
                r = ref[k]
                errs[k] = float((got[k] - r).abs().max() / r.abs().max().clamp_min(1e-30))
            worst = torch.tensor([max(errs.values())], dtype=torch.float64, device=dev)
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            ms = timed(step_compact)
            # results must still be right after many back-to-back steps (double-buffered slots, epoch counters)
            got = step_compact()
            torch.cuda.synchronize()
            for k in ("dmeans3D", "dcov3D", "dopacity", "dsh"):
                errs[k] = max(errs[k], float((got[k] - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30)))
            worst = torch.tensor([max(errs.values())], dtype=torch.float64, device=dev)
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            phases = {}
            if "graph" not in transport:
                ex.profile = True
                for _ in range(10):
                    flush_buf.zero_()
                    step_compact()
                    for k, v in ex.phase_ms().items():
                        phases[k] = phases.get(k, 0.0) + v / 10
                ex.profile = False
            if not worst.item() < 1e-4:
                failures.append(transport)
            say({"variant": f"compact_{transport}", "world": world, "ms_per_step": ms, "max_rel_err_vs_arena": errs,
                 "worst_over_ranks": float(worst.item()), "ok": bool(worst.item() < 1e-4),
                 "multicast": bool(ex.handles and ex.handles[1].multicast_ptr), "bytes": ex.exchange_bytes(),
                 "phase_ms_rank0": {k: round(v, 4) for k, v in phases.items()}})
        except Exception as e:  # report and carry on with the next variant
            failures.append(transport)
            say({"variant": f"compact_{transport}", "world": world, "error": repr(e)[:400],
                 "trace": traceback.format_exc()[-1200:]})
            torch.cuda.synchronize()
    dist.destroy_process_group()
    if args.check and failures:
        raise SystemExit(f"multi_gpu_check: failed variants {failures}")


if __name__ == "__main__":
    main()
