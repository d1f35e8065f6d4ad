#!/usr/bin/env python
"""Where does the full-size (config #2) parity error come from?  Same weights / inputs as bench.py (seed 0), B samples:
CPU oracle (fp32) vs native `tc` vs native `simt` (exact-fp32 FFMA convs) vs the torch formulation through cuDNN fp32 on the GPU.
Prints the parity metrics (max|a-b| / max(|b|, std b)) of every pair: the spread between the three fp32 paths is the noise
floor of the problem (summation order through 152 + 33 layers); `tc` should sit within a small factor of it."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lt_b200  # noqa: E402
from lt_b200 import testing  # noqa: E402
from oracle import parity  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--layers", type=int, default=152)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, V, S, n = a.batch, 4, 384, 64
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
np.random.seed(0)
cfg = lambda: testing.make_config(num_layers=a.layers, volume_size=n)   # noqa: E731
holder = lt_b200.VolumetricTriangulationNet(cfg(), device=dev, backend="torch")
testing.randomize_weights(holder, seed=a.seed, calib_size=S, calib_views=1)
holder = holder.to(dev).eval()
sd = {k: v.detach().cpu() for k, v in holder.state_dict().items()}
images, batch = testing.make_batch(B, V, image_size=S, seed=a.seed)
torch.set_num_threads(min(os.cpu_count(), 32))
oracle_out, secs = parity.oracle_forward(sd, images, batch, n)
outs = {}
with torch.no_grad():
    outs["cudnn_fp32"] = [t.detach().cpu() if torch.is_tensor(t) else t for t in holder(images.to(dev), None, batch)]
del holder
torch.cuda.empty_cache()
for mode in ("tc", "simt"):
    m = lt_b200.VolumetricTriangulationNet(cfg(), device="cpu", backend="native", conv_mode=mode, use_cuda_graph=False)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    with torch.no_grad():
        outs[mode] = [t.detach().cpu() if torch.is_tensor(t) else t for t in m(images.to(dev), None, batch)]
    del m
    torch.cuda.empty_cache()


def cmp(x, ref):      # x, ref: 7-tuples (kp, feats, vols, ...)
    kp, f, v = x[0], x[1], x[2]
    kp_r, f_r, v_r = ref[0], ref[1], ref[2]
    J = v.shape[1]
    return {"features": parity.rel_err(f.numpy(), f_r.numpy()), "volumes": parity.rel_err(v.numpy(), v_r.numpy()),
            "keypoints_mm": float((kp - kp_r).abs().max()),
            "argmax_equal": bool(torch.equal(v.reshape(B, J, -1).argmax(-1), v_r.reshape(B, J, -1).argmax(-1)))}


ref = (oracle_out[0], oracle_out[1], oracle_out[2])
res = {"oracle_seconds": secs, "batch": B}
for k in outs:
    res[k + " vs cpu_oracle"] = cmp(outs[k], ref)
res["tc vs simt"] = cmp(outs["tc"], outs["simt"])
res["tc vs cudnn_fp32"] = cmp(outs["tc"], outs["cudnn_fp32"])
res["simt vs cudnn_fp32"] = cmp(outs["simt"], outs["cudnn_fp32"])
print(json.dumps(res))
