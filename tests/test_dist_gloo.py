"""world_size-2 (and 4) gloo tests of the view-sharding algebra and orchestration (dist.py), CPU only.

Each rank computes the partial aggregate of ITS views with the torch formulation, the group completes it with one
collective, and the result must equal the single-process aggregation over all views (softmax: packed num/den)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lt_b200 import dist as lt_dist, testing, torch_ops
from oracle import vol_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene(B, V, C, h, w, n):
    rng = np.random.RandomState(5)
    heat = torch.from_numpy(rng.randn(B, V, C, h, w).astype(np.float32))
    cams = testing.make_cameras(V, image_size=48, radius=3000.0)
    proj = torch.from_numpy(np.stack([np.stack([O.projection_after_resize(c.K, c.R, c.t, (48, 48), (h, w)) for c in cams])] * B))
    coord = torch.from_numpy(np.stack([O.coord_volume(rng.randn(3) * 100 + [0, 0, 900], 2600.0, n) for _ in range(B)]))
    return heat, proj, coord


def _worker(rank, world, port, n_views, agg, collective, ret, max_bytes=1 << 29):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = lt_dist.make_plan(world, rank, n_views)
        pg = lt_dist.new_view_groups(plan)
        B = 2 * plan.group_size
        heat, proj, coord = _scene(B, n_views, 4, 9, 11, 5)
        # different data-parallel groups work on different samples: perturb by group index
        heat = heat + 0.1 * plan.group_index
        vs = plan.views
        sampled = torch_ops.sample_views(heat[:, vs], proj[:, vs], coord)
        partial = torch_ops.partial_aggregate(sampled, agg)
        mine = lt_dist.complete_partials(partial, plan, pg, collective, "max" if agg == "max" else "sum", max_bytes=max_bytes)
        vol_local = torch_ops.finalize_aggregate(mine, agg)
        full = torch_ops.unproject_heatmaps(heat, proj, coord, agg).reshape(B, 4, -1)
        own = plan.owned_samples(B)
        err = float((vol_local - full[own[0]:own[-1] + 1]).abs().max() / full.abs().max())
        kp_local = vol_local.mean(dim=2)[:, :3].unsqueeze(1).contiguous()     # stand-in (B/G, 1, 3) "keypoints"
        kp_all = lt_dist.gather_keypoints(kp_local, plan, pg)
        err_g = float((kp_all - full.mean(dim=2)[:, :3].unsqueeze(1)).abs().max())
        ret[rank] = (err, err_g, plan.group_size, plan.n_groups, len(vs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_views,agg", [(2, 4, "softmax"), (2, 4, "sum"), (2, 4, "max"), (2, 3, "softmax"), (4, 4, "softmax")])
def test_view_sharded_aggregation_equals_single_process(world, n_views, agg):
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n_views, agg, "all_reduce", ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank, (err, err_g, g, ng, nv) in ret.items():
        assert g * ng == world and nv == n_views // g
        # the unshifted exp num/den partials, summed and divided, differ from torch.softmax at fp32 rounding level
        tol = 1e-4 if agg == "softmax" else 1e-5
        assert err < tol and err_g < tol, (rank, err, err_g)


def test_all_reduce_in_slices_equals_one_call():
    """The packed partials are all-reduced in slices of whole samples (dist.complete_partials: at most 512 MiB per call on the GPU
    path); forcing one sample per call must give the same aggregate."""
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), 4, "softmax", "all_reduce", ret, 64), nprocs=2, join=True)
    assert len(ret) == 2
    for rank, (err, err_g, g, ng, nv) in ret.items():
        assert err < 1e-4 and err_g < 1e-4, (rank, err, err_g)


def test_plan_partitions_views_and_samples():
    for world, views in [(1, 4), (2, 4), (4, 4), (8, 4), (8, 8), (2, 3), (4, 6)]:
        plans = [lt_dist.make_plan(world, r, views) for r in range(world)]
        g = plans[0].group_size
        assert world % g == 0 and views % g == 0
        for gi in range(plans[0].n_groups):
            members = [p for p in plans if p.group_index == gi]
            assert sorted(v for p in members for v in p.views) == list(range(views))
            assert sorted(s for p in members for s in p.owned_samples(2 * g)) == list(range(2 * g))


def _feature_worker(rank, world, port, n_views, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = lt_dist.make_plan(world, rank, n_views)
        pg = lt_dist.new_view_groups(plan)
        assert plan.n_groups > 1, "this test covers the several-groups (NCCL / gloo all-gather) variant of the feature exchange"
        per, h, w, C = 3, 4, 5, 8
        B = per * plan.group_size
        g = torch.Generator().manual_seed(7 + plan.group_index)
        feats_all = torch.randn(B, n_views, h, w, C, generator=g)             # the group's true maps, all views
        fx = lt_dist.FeatureExchange(plan, pg, B, n_views, h, w, C, torch.device("cpu"))
        fx.barrier()
        fx.scatter(feats_all[:, plan.views].contiguous())
        fx.barrier()
        own = plan.owned_samples(B)
        ret[rank] = bool(torch.equal(fx.buf, feats_all[own[0]:own[-1] + 1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_views", [(4, 2), (4, 6)])
def test_feature_exchange_with_several_view_groups(world, n_views):
    """world 4, 2 views -> 2 groups x 2 ranks (one view each); 6 views -> 2 groups x 2 ranks with 3 views per rank: every owner must end
    up with exactly its samples' maps of ALL views, in global view order (the layout lt_unproject_aggregate_fwd reads)."""
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_feature_worker, args=(world, _free_port(), n_views, ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret.values()), dict(ret)
