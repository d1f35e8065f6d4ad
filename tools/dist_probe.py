#!/usr/bin/env python
"""Diagnose the view-sharded exchanges on N GPUs (torchrun): per-rank key-point error of every exchange against a single-GPU forward of
the rank's own samples (eager and graph-captured), and -- for the `features` exchange -- the owner's assembled peer buffer against the
true feature maps gathered with NCCL, per (sample, view).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/dist_probe.py [--views 4]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lt_b200  # noqa: E402
from lt_b200 import dist as lt_dist, testing  # noqa: E402
from lt_b200.triangulation import backbone_map_size  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--batch", type=int, default=2, help="samples per rank")
ap.add_argument("--layers", type=int, default=50)
ap.add_argument("--image", type=int, default=256)
ap.add_argument("--volume", type=int, default=32)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = testing.make_config(num_layers=a.layers, volume_size=a.volume)
torch.manual_seed(0)
np.random.seed(0)
model = lt_b200.VolumetricTriangulationNet(cfg, device=dev, backend="native", conv_mode="tc", use_cuda_graph=False)
testing.randomize_weights(model, seed=0, calib_size=a.image, calib_views=1)
model = model.to(dev).eval()
eng = model.engine()
plan = lt_dist.make_plan(world, rank, a.views)
pg = lt_dist.new_view_groups(plan)
Bg = a.batch * plan.group_size
images_g, batch = testing.make_batch(Bg, a.views, image_size=a.image, seed=100 + plan.group_index)
images = images_g[:, plan.views].contiguous().to(dev)
hm = (backbone_map_size(a.image),) * 2
proj, base, position, stepv, rots, _ = model._host_geometry(batch, Bg, (a.image, a.image), hm)
up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)   # noqa: E731
geo = (up(proj[:, plan.views]), up(position), up(base), up(stepv), up(rots.reshape(Bg, 9)), up(proj))
own = plan.owned_samples(Bg)
lo, hi = own[0], own[-1] + 1
sub = {"cameras": [c[lo:hi] for c in batch["cameras"]], "keypoints_3d": batch["keypoints_3d"][lo:hi],
       "pred_keypoints_3d": batch["pred_keypoints_3d"][lo:hi]}
with torch.no_grad():
    kp_single = model(images_g[lo:hi].to(dev), None, sub)[0]
    lines = []
    for coll in ("all_reduce", "reduce_scatter", "p2p", "features"):
        for graph in (False, True):
            try:
                kp = eng.forward_view_sharded(images, geo[0], geo[1], geo[2], geo[3], geo[4], plan, pg, coll, proj_all=geo[5], use_graph=graph)[0]
                if graph:   # replay once more: the captured path, not the capture-time eager pass
                    kp = eng.forward_view_sharded(images, geo[0], geo[1], geo[2], geo[3], geo[4], plan, pg, coll, proj_all=geo[5], use_graph=True)[0]
                torch.cuda.synchronize()
                err = (kp[lo:hi] - kp_single).abs().amax(dim=(1, 2))
                lines.append("%s%s: max %.4f mm per own sample %s" % (coll, "+graph" if graph else "", float(err.max()), [round(float(e), 3) for e in err]))
            except Exception as exc:   # noqa: BLE001
                lines.append("%s%s: FAILED %s: %s" % (coll, "+graph" if graph else "", type(exc).__name__, str(exc)[:200]))
    # features exchange: owner's buffer vs the truth
    eng.forward_view_sharded(images, geo[0], geo[1], geo[2], geo[3], geo[4], plan, pg, "features", proj_all=geo[5], use_graph=False)
    torch.cuda.synchronize()
    feats = eng.backbone_features(images.reshape(Bg * len(plan.views), *images.shape[2:]))
    mine = feats.data.view(Bg, len(plan.views), feats.H, feats.W, feats.C).contiguous()
    allf = [torch.empty_like(mine) for _ in range(plan.group_size)]
    dist.all_gather(allf, mine, group=pg)
    torch.cuda.synchronize()
    buf = eng._peer.buf            # (per, V, h, w, C)
    bad = []
    for bl, b in enumerate(range(lo, hi)):
        for v in range(a.views):
            src_rank, j = v % plan.group_size, v // plan.group_size
            want = allf[src_rank][b, j]
            d = float((buf[bl, v] - want).abs().max())
            if d != 0.0:
                bad.append((b, v, round(d, 5)))
    lines.append("features peer buffer vs gathered truth: %d mismatching (sample, view) of %d: %s" % (len(bad), (hi - lo) * a.views, bad[:12]))
for r in range(world):
    dist.barrier()
    if r == rank:
        print("== rank %d (group %d, view_rank %d, views %s, own samples %d..%d)" % (rank, plan.group_index, plan.view_rank, plan.views, lo, hi - 1))
        print("\n".join("   " + l for l in lines), flush=True)
dist.destroy_process_group()
