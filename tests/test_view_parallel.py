"""World-size-2 gloo test of the view-sharded path (CPU; the rasterizer is the oracle backend).

Rank r renders views r, r+2 of the same Gaussians; the all-reduced gradient arena must equal
the single-process sum over all views (SURVEY.md 8e verification rule)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ggrt_official_b200.render import render_cuda
from ggrt_official_b200.synthetic import make_scene, small_se3
from ggrt_official_b200.view_parallel import all_reduce_gradients, render_views_sharded, shard_views

VIEWS, P, H, W = 4, 600, 32, 48


def _scene():
    sc = make_scene(P, H, W, sh_degree=2, seed=21)
    rng = np.random.default_rng(3)
    extr = [sc.extrinsics.astype(np.float64)]
    for _ in range(VIEWS - 1):
        extr.append(extr[0] @ small_se3(rng).astype(np.float64))
    return sc, torch.tensor(np.stack(extr).astype(np.float32))


def _leaves(sc):
    t = torch.tensor
    return [t(sc.means)[None].requires_grad_(), t(sc.covariances)[None].requires_grad_(),
            t(sc.harmonics)[None].requires_grad_(), t(sc.opacities)[None].requires_grad_()]


def _render_view(v, sc, extr, leaves, weights):
    means, cov, harm, opac = leaves
    img = render_cuda(extr[v][None], torch.tensor(sc.intrinsics)[None], torch.tensor([sc.near]),
                      torch.tensor([sc.far]), (H, W), torch.zeros(1, 3), means, cov, harm, opac)
    return (img * weights[v]).sum()


def _worker(rank, world, port, out):
    from tests import oracle_backend as ob

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc, extr = _scene()
        leaves = _leaves(sc)
        weights = torch.randn(VIEWS, 1, 3, H, W, generator=torch.Generator().manual_seed(9))
        with ob.installed():
            losses = render_views_sharded(lambda v: _render_view(v, sc, extr, leaves, weights), VIEWS)
            assert sorted(losses) == shard_views(VIEWS, rank, world)
            sum(losses.values()).backward()
        all_reduce_gradients(leaves)
        if rank == 0:
            torch.save([l.grad for l in leaves], out)
    finally:
        dist.destroy_process_group()


def test_two_rank_view_sharding_matches_single_process(tmp_path):
    from tests import oracle_backend as ob

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)

    sc, extr = _scene()
    leaves = _leaves(sc)
    weights = torch.randn(VIEWS, 1, 3, H, W, generator=torch.Generator().manual_seed(9))
    with ob.installed():
        sum(_render_view(v, sc, extr, leaves, weights) for v in range(VIEWS)).backward()
    for g, l in zip(got, leaves):
        ref = l.grad
        assert float((g - ref).abs().max()) <= 1e-3 * float(ref.abs().max()) + 1e-12


# ---- compact exchange: gather [P,3] colour gradients + all-reduce [P,10], rebuild dL/dsh locally -------------
def _raster_inputs(view):
    from ggrt_official_b200.synthetic import to_raster_inputs

    sc, extr = _scene()
    sc.extrinsics = extr[view].numpy()
    return to_raster_inputs(sc, bg=(0.1, 0.0, 0.2))


def _settings(ri):
    from ggrt_official_b200 import GaussianRasterizationSettings

    t = torch.tensor
    return GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=ri.tanfovx, tanfovy=ri.tanfovy, bg=t(ri.bg), scale_modifier=1.0,
        viewmatrix=t(ri.viewmatrix), projmatrix=t(ri.projmatrix), sh_degree=ri.sh_degree, campos=t(ri.campos),
        prefiltered=False)


def _compact_worker(rank, world, port, out):
    from ggrt_official_b200.view_parallel import CompactGradientExchange
    from tests import oracle_backend as ob

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ri = _raster_inputs(rank)
        t = torch.tensor
        st = ob.fake_forward_raw(t(ri.means3D), t(ri.shs), None, t(ri.opacities), t(ri.cov3D), _settings(ri))
        ex = CompactGradientExchange(P, ri.sh_degree, "cpu", backward_fn=ob.fake_backward_compact,
                                     merge_fn=ob.fake_sh_gradient_merge)
        assert ex.exchange_bytes()["gather_recv"] == 4 * 3 * (P + 1) * (world - 1)
        w = torch.randn(VIEWS, 3, H, W, generator=torch.Generator().manual_seed(9))
        g = ex.run(st, w[rank])
        if rank == 0:
            torch.save({k: v.clone() for k, v in g.items() if k != "dmeans2D"}, out)
    finally:
        dist.destroy_process_group()


def test_compact_exchange_matches_sum_of_full_gradients(tmp_path):
    from oracle import c_oracle as co
    from tests.helpers import oracle_camera

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "compact.pt")
    mp.spawn(_compact_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)

    w = torch.randn(VIEWS, 3, H, W, generator=torch.Generator().manual_seed(9)).numpy()
    ref = None
    for v in range(2):
        ri = _raster_inputs(v)
        cam = oracle_camera(ri)
        f = co.forward(cam, ri.means3D, ri.cov3D, ri.opacities, sh=ri.shs)
        b = co.backward(cam, ri.means3D, ri.cov3D, ri.opacities, f, w[v], sh=ri.shs)
        cur = dict(dmeans3D=b["dmeans3D"], dcov3D=b["dcov3D"], dopacity=b["dopacity"][:, None], dsh=b["dsh"])
        ref = cur if ref is None else {k: ref[k] + cur[k] for k in ref}
    for k, r in ref.items():
        g = got[k].numpy()
        assert np.abs(r).max() > 0
        assert np.abs(g - r).max() <= 1e-4 * np.abs(r).max(), k
