#!/usr/bin/env python
"""One eager (graph-free) step of the native path at BASELINE config #2 shapes, for ncu.

    ncu ... python tools/profile_step.py [--stage all|v2v|post] [--mode tc] [--batch 8]
Default-init weights (timing only).  `--stage v2v` runs only unprojection + V2V + soft-argmax on random features.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lt_b200  # noqa: E402
from lt_b200 import capi, testing  # noqa: E402
from lt_b200.engine import Act  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--stage", default="all")
ap.add_argument("--mode", default="tc")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--repeat", type=int, default=2)
a = ap.parse_args()

dev = "cuda:0"
B, V, S, n = a.batch, a.views, 384, 64
model = lt_b200.VolumetricTriangulationNet(testing.make_config(num_layers=152, volume_size=n), device=dev, backend="native",
                                           conv_mode=a.mode, use_cuda_graph=False).to(dev).eval()
images, batch = testing.make_batch(B, V, image_size=S, seed=0)
images = images.to(dev)
eng = model.engine()
with torch.no_grad():
    if a.stage == "all":
        for _ in range(a.repeat):
            model(images, None, batch)
    else:
        eng.prepare()
        proj, base, position, step, rots, _ = model._host_geometry(batch, B, (S, S), (S // 4, S // 4))
        up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
        coord = torch.empty((B, n, n, n, 3), device=dev)
        capi.coord_volume(up(position), up(base), up(step), up(rots.reshape(B, 9)), coord, False)
        feats = Act(B * V, 1, S // 4, S // 4, 32, capi.FMT_F32, dev)
        feats.data.normal_()
        for _ in range(a.repeat):
            vol = eng.unproject(feats, B, V, up(proj), coord, capi.AGG["softmax"])
            if a.stage == "v2v":
                logits = eng.v2v(vol, (coord, 17, 1.0, True))      # fused tail: carries the soft-argmax statistics pass
                eng.softargmax(logits, coord, 17, 1.0, True)
torch.cuda.synchronize()
print("done", eng.launches)
