#!/usr/bin/env python
"""One TRAINING step (train.py:189-243: forward in train mode, L1 key-point loss, backward, Adam step) of the volumetric model on
one B200, timed for backend="torch" (the autograd formulation of the whole path through ATen/cuDNN) and backend="hybrid"
(same convolutions, but the unprojection + aggregation and the soft-argmax run on the native forward AND backward kernels,
csrc/unproject.cu + csrc/backward.cu).

    python tools/train_step_bench.py [--batch 4] [--layers 152] [--steps 5]
Prints one JSON line: ms per step and peak memory for both backends, and the gradient agreement of the first step.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lt_b200  # noqa: E402
from lt_b200 import testing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--image", type=int, default=384)
ap.add_argument("--volume", type=int, default=64)
ap.add_argument("--layers", type=int, default=152)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False          # true fp32 convolutions in both arms
torch.backends.cuda.matmul.allow_tf32 = False

images, batch = testing.make_batch(a.batch, a.views, image_size=a.image, seed=0)
images = images.to(dev)
gt = torch.from_numpy(np.stack(batch["keypoints_3d"])[:, :, :3]).float().to(dev)
out = {}
grads = {}
sd0 = None
for backend in ("torch", "hybrid"):
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = testing.make_config(num_layers=a.layers, volume_size=a.volume)
    model = lt_b200.VolumetricTriangulationNet(cfg, device=dev, backend=backend)
    if sd0 is None:
        testing.randomize_weights(model, seed=0, calib_size=a.image, calib_views=1)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    else:
        model.load_state_dict(sd0)
    model = model.to(dev).train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-5)
    torch.cuda.reset_peak_memory_stats()

    def step():
        np.random.seed(1)                                   # same random cuboid rotation for both backends
        kp = model(images, None, batch)[0]
        loss = (kp - gt).abs().mean()                       # KeypointsMAELoss with full validity (loss.py:20-28)
        opt.zero_grad()
        loss.backward()
        return loss

    loss0 = step()
    grads[backend] = model.process_features[0].weight.grad.detach().clone()
    opt.step()
    for _ in range(2):
        step(); opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step(); opt.step()
    e1.record()
    torch.cuda.synchronize()
    out[backend] = {"ms_per_step": e0.elapsed_time(e1) / a.steps, "samples_per_s": a.batch * a.steps / (e0.elapsed_time(e1) / 1e3),
                    "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "first_loss": float(loss0)}
    del model, opt
    torch.cuda.empty_cache()
g_t, g_h = grads["torch"], grads["hybrid"]
print(json.dumps({"what": "training step, ResNet-%d, %d views %dx%d, %d^3 grid, batch %d, fp32 (cuDNN convs), Adam" % (a.layers, a.views, a.image, a.image, a.volume, a.batch),
                  "torch": out["torch"], "hybrid": out["hybrid"], "speedup_hybrid_over_torch": out["torch"]["ms_per_step"] / out["hybrid"]["ms_per_step"],
                  "first_step_grad_rel_diff(process_features.weight)": float((g_t - g_h).abs().max() / g_t.abs().max())}))
