"""2-GPU test of the view-sharded path (needs >= 2 CUDA devices): NCCL all-reduce, reduce-scatter and the fused
P2P-store exchange must all reproduce the single-GPU result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    import lt_b200
    from lt_b200 import testing, dist as lt_dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, V, S, n = 2, 2, 128, 32
        cfg = testing.make_config(num_layers=50, volume_size=n)
        holder = lt_b200.VolumetricTriangulationNet(cfg, device="cpu", backend="torch")
        testing.randomize_weights(holder, seed=1, calib_size=S)
        model = lt_b200.VolumetricTriangulationNet(testing.make_config(num_layers=50, volume_size=n), device=dev, backend="native",
                                                   conv_mode="tc", use_cuda_graph=False)
        model.load_state_dict(holder.state_dict())
        model = model.to(dev).eval()
        images, batch = testing.make_batch(B, V, image_size=S, seed=3)
        with torch.no_grad():
            kp_single = model(images.to(dev), None, batch)[0]
            plan = lt_dist.make_plan(world, rank, V)
            pg = lt_dist.new_view_groups(plan)
            from lt_b200.triangulation import backbone_map_size
            proj, base, position, step, rots, _ = model._host_geometry(batch, B, (S, S), (backbone_map_size(S),) * 2)
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
            vs = plan.views
            errs = {}
            for coll in ("all_reduce", "reduce_scatter", "p2p", "features"):
                for graph in (False, True):      # graph: stages 1 and 3 replayed as CUDA graphs around the eager exchange
                    key = coll + ("+graph" if graph else "")
                    try:
                        for _ in range(2 if graph else 1):   # second call replays the captured graphs
                            kp = model.engine().forward_view_sharded(images[:, vs].contiguous().to(dev), up(proj[:, vs]), up(position), up(base),
                                                                     up(step), up(rots.reshape(B, 9)), plan, pg, coll, proj_all=up(proj),
                                                                     use_graph=graph)[0]
                        torch.cuda.synchronize()
                        errs[key] = float((kp - kp_single).abs().max())
                    except Exception as e:   # report, do not hang the other rank
                        errs[key] = "ERROR: %r" % (e,)
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


def test_view_sharded_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    print(dict(ret))
    for rank, errs in ret.items():
        for coll, e in errs.items():
            assert not isinstance(e, str), (rank, coll, e)
            assert e < 0.05, (rank, coll, e)     # mm; partial sums are reordered, nothing else changes
