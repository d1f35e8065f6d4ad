#!/usr/bin/env python
"""Single-GPU check of the view-sharded all_reduce arithmetic at the sizes of a 32-sample view group (packed partials = 2 GiB):
per-view partials (lt_unproject_partial_fwd, one view at a time as the four ranks of a group would compute them) summed with torch,
finalised per 8-sample block (lt_unproject_finalize_fwd) and compared with the fused single-GPU kernel (lt_unproject_aggregate_fwd)
on the same inputs.  No NCCL, no CUDA graphs: if this agrees, the kernels are not what breaks at N = 4 / 8."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lt_b200  # noqa: E402
from lt_b200 import capi, testing  # noqa: E402

dev = "cuda:0"
V, h, C, n = 4, 96, 32, 64
nvox = n ** 3
model = lt_b200.VolumetricTriangulationNet(testing.make_config(num_layers=50, volume_size=n), device=dev, backend="native", use_cuda_graph=False)
for B in (16, 32):
    _, batch = testing.make_batch(B, V, image_size=384, seed=100)
    proj, base, position, step, rots, _ = model._host_geometry(batch, B, (384, 384), (h, h))
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)   # noqa: E731
    coord = torch.empty((B, n, n, n, 3), device=dev)
    capi.coord_volume(up(position), up(base), up(step), up(rots.reshape(B, 9)), coord, False)
    g = torch.Generator(device=dev).manual_seed(B)
    feats = torch.randn(B, V, h, h, C, device=dev, generator=g) * 2.0
    projd = up(proj)
    full = torch.empty((B, nvox, C), dtype=torch.float32, device=dev)
    capi.unproject_aggregate(feats, projd, coord.view(B, nvox, 3), None, full, capi.FMT_F32, capi.AGG["softmax"])
    parts = torch.zeros((B, 2, nvox, C), dtype=torch.float32, device=dev)
    tmp = torch.empty_like(parts)
    for v in range(V):
        capi.unproject_partial(feats[:, v:v + 1].contiguous(), projd[:, v:v + 1].contiguous(), coord.view(B, nvox, 3), None, tmp, capi.AGG["softmax"])
        parts += tmp
    worst = 0.0
    for r in range(B // 8):
        out = torch.empty((8, nvox, C), dtype=torch.float32, device=dev)
        capi.unproject_finalize(parts[8 * r:8 * r + 8].contiguous(), out, capi.FMT_F32, 8, C, nvox, capi.AGG["softmax"])
        torch.cuda.synchronize()
        d = float((out - full[8 * r:8 * r + 8]).abs().max()) / float(full.abs().max())
        worst = max(worst, d)
        print("B %d block %d: max |partial path - fused| / max|fused| = %.3e" % (B, r, d), flush=True)
    print("B %d (packed partials %.2f GiB): worst %.3e" % (B, parts.numel() * 4 / 2 ** 30, worst), flush=True)
    del parts, tmp, full, feats, coord
    torch.cuda.empty_cache()
