"""Multi-GPU hardware test of the view-sharded gradient exchange (needs >= 2 GPUs; skipped otherwise).

Launches tools/multi_gpu_check.py under torchrun with one rank per GPU: every rank renders another target view of
the same Gaussians; the compact exchange on every transport (NCCL; symmetric-memory push with in-kernel signalling,
eager and as one captured CUDA graph; with torch's / this library's barrier kernels) must reproduce the NCCL
all-reduce of the full gradient arena within float32 summation noise, also after many back-to-back steps."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_compact_exchange_matches_arena_allreduce_on_2_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", str(ROOT / "tools" / "multi_gpu_check.py"), "--workload", "c1",
           "--gaussians", "60000", "--iters", "5", "--check"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    rows = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-2000:])
    names = {r["variant"] for r in rows}
    assert {"arena_allreduce", "compact_nccl", "compact_p2p+signal", "compact_p2p+signal+graph"} <= names, names
    for r in rows:
        if r["variant"].startswith("compact_"):
            assert r.get("ok"), r
