"""GPU tests of graph.CapturedStep: forward + backward frozen into one CUDA graph (programmatic dependent launch along
the kernel chain, backward scratch zeroed during the forward, dL/dsh writer beside the per-Gaussian kernel) against the
eagerly launched calls -- images bit for bit, gradients up to the summation order of the float atomics -- over several
replays, with changed inputs, and on scenes that reach the multi-warp sort tiers."""
import numpy as np
import pytest
import torch

from ggrt_official_b200 import rasterizer as R
from ggrt_official_b200.graph import CapturedStep
from tests import gpu_util as G
from tests.helpers import small_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, tol=3e-5):
    return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("P,H,W,deg,cs,seed,aux,pose", [
    (3000, 100, 75, 4, 9.0, 2, False, False),
    (1400, 32, 32, 1, 60.0, 7, False, False),   # one dense tile row: 8-warp sort tier, multi-batch render kernels
    (4000, 32, 32, 0, 60.0, 8, True, False),    # degree 0 (no colour Jacobian), aux channel with gradient
    (20000, 160, 208, 3, 2.0, 5, False, True),  # camera gradients
])
def test_captured_step_equals_the_eager_calls(P, H, W, deg, cs, seed, aux, pose):
    _, ri = small_case(P, H, W, deg, seed=seed, cov_scale=cs)
    t = lambda a: torch.tensor(np.asarray(a), device=DEV)
    rs = G.settings_from(ri, DEV)
    means, sh, opac, cov = t(ri.means3D), t(ri.shs), t(ri.opacities), t(ri.cov3D)
    auxv = torch.rand(P, device=DEV) if aux else None
    rng = np.random.default_rng(seed)
    g = t(rng.standard_normal((3, H, W)).astype(np.float32))
    ga = t(rng.standard_normal((H, W)).astype(np.float32)) if aux else None

    def eager():
        st = R.forward_raw(means, sh, None, opac, cov, rs, aux=auxv)
        return st, R.backward_raw(st, g, grad_aux=ga, want_camera=pose)

    cap = CapturedStep(means, sh, None, opac, cov, rs, grad_color=g, aux=auxv, grad_aux=ga, want_camera=pose)
    for trial in range(3):
        if trial == 1:  # new loss gradient, written in place
            g.copy_(t(rng.standard_normal((3, H, W)).astype(np.float32)))
        if trial == 2:  # new Gaussians, written in place (same shape; the pair count changes a little)
            means.add_(0.01 * torch.randn_like(means))
            opac.mul_(0.9)
        for _ in range(3):  # a scratch that was not zeroed again would show up as doubled gradients
            cap.replay()
        N, _ = cap.check()
        st, ref = eager()
        assert N == st["N"]
        assert torch.equal(cap.color, st["color"]) and torch.equal(cap.depth, st["depth"]) and torch.equal(cap.radii, st["radii"])
        for k, v in ref.items():
            if v is None:
                continue
            # (the 35 camera sums go through a few thousand float atomics: looser)
            assert _close(cap.grads[k], v, 1e-3 if k == "dcamera" else 3e-5), (trial, k, float((cap.grads[k] - v).abs().max()), float(v.abs().max()))


@pytest.mark.parametrize("env", [{"GGRT_RASTER_OVERLAP": "0"}, {"GGRT_RASTER_PDL": "0"}])
def test_library_switches_keep_the_results(env):
    """GGRT_RASTER_OVERLAP=0 (every kernel and the scratch memset on the caller's stream) and GGRT_RASTER_PDL=0 (no
    programmatic dependent launch under capture) are read once per process: the captured-step and oracle-parity tests
    are re-run in a child process with the switch set."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x", "tests/test_gpu_graph.py", "-k",
                          "(captured_step and (3000 or 20000)) or test_gpu_parity", "tests/test_gpu_parity.py"], cwd=root, capture_output=True,
                         text=True, timeout=900, env=dict(os.environ, **env))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert " passed" in res.stdout
