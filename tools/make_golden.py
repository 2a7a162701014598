"""Generates tests/golden/decoder_*.npz by running the UNMODIFIED reference decoder
(/root/reference/ggrt/model/pixelsplat/decoder/decoder_splatting_cuda.py -> cuda_splatting.py)
on CPU with the oracle standing in for the CUDA rasterizer (tests/oracle_backend.py).

The fixtures pin the caller-side semantics (camera conventions, scale-invariant rescale, SH
layout, depth-as-colour trick, per-view loop) against the real reference code; the
rasterizer-side numbers come from the oracle (PARITY UNPINNED, see oracle/raster_oracle.c).
Run in the build container only (needs /root/reference):  python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ggrt_official_b200.synthetic import make_scene, small_se3  # noqa: E402
from tests import oracle_backend as ob  # noqa: E402
from tests.ref_import import load_reference_glue  # noqa: E402

CASES = {
    # name: (P, H, W, sh_degree, views, near, seed, depth_mode)
    "decoder_a": (1500, 64, 80, 4, 1, 1.0, 11, "depth"),
    "decoder_b": (1200, 48, 64, 4, 2, 0.5, 12, "depth"),   # two views, near != 1 exercises the rescale
    "decoder_c": (800, 40, 40, 4, 1, 2.0, 13, "disparity"),
}


def decoder_inputs(P, H, W, deg, views, near, seed):
    sc = make_scene(P, H, W, sh_degree=deg, seed=seed, near=1.0, far=100.0)
    k = np.float32(near)  # scale the whole scene so that the near plane sits at `near`
    rng = np.random.default_rng(seed + 7)
    extr = [sc.extrinsics.astype(np.float64)]
    for _ in range(views - 1):
        extr.append(extr[0] @ small_se3(rng).astype(np.float64))
    extr = np.stack(extr).astype(np.float32)
    extr[:, :3, 3] *= k
    return dict(
        means=(sc.means * k)[None], covariances=(sc.covariances * k * k)[None], harmonics=sc.harmonics[None],
        opacities=sc.opacities[None], extrinsics=extr[None], intrinsics=np.repeat(sc.intrinsics[None], views, 0)[None],
        near=np.full((1, views), near, np.float32), far=np.full((1, views), 100.0 * near, np.float32),
        image_shape=np.asarray([H, W]))


def run_reference(inp, depth_mode):
    cs, dec = load_reference_glue()
    from ggrt.model.pixelsplat.types import Gaussians

    t = {k: torch.tensor(v) for k, v in inp.items() if k != "image_shape"}
    leaves = {k: t[k].clone().requires_grad_() for k in ("means", "covariances", "harmonics", "opacities")}
    decoder = dec.DecoderSplattingCUDA(dec.DecoderSplattingCUDACfg(name="splatting_cuda"))
    with ob.installed():
        out = decoder.forward(Gaussians(**leaves), t["extrinsics"], t["intrinsics"], t["near"], t["far"],
                              tuple(int(x) for x in inp["image_shape"]), depth_mode=depth_mode)
        g = torch.Generator().manual_seed(5)
        wc = torch.randn(out.color.shape, generator=g)
        wd = torch.randn(out.depth.shape, generator=g) * 0.1
        ((out.color * wc).sum() + (out.depth * wd).sum()).backward()
    res = dict(color=out.color.detach().numpy(), depth=out.depth.detach().numpy(), w_color=wc.numpy(), w_depth=wd.numpy())
    for k, v in leaves.items():
        res["grad_" + k] = v.grad.numpy()
    return res


if __name__ == "__main__":
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(exist_ok=True)
    for name, (P, H, W, deg, views, near, seed, mode) in CASES.items():
        inp = decoder_inputs(P, H, W, deg, views, near, seed)
        res = run_reference(inp, mode)
        np.savez_compressed(out_dir / f"{name}.npz", depth_mode=mode, **{"in_" + k: v for k, v in inp.items()},
                            **{"out_" + k: v.astype(np.float32) for k, v in res.items()})
        print(name, {k: v.shape for k, v in res.items()}, "color range", res["color"].min(), res["color"].max(),
              "depth max", res["depth"].max())
