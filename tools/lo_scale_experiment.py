"""CPU emulation of the split-fp16 product schemes on REAL activations of the calibrated ResNet-50 / V2V model:
per conv layer, the relative error (max|d| / max(|y|, std y) against an fp64 convolution) of
  scaled   x = hi + lo/2048, lo = fp16((x - hi) * 2048): today's format, cross terms in a second accumulator
  unscaled x = hi + lo,      lo = fp16(x - hi)           : one accumulator (fp16 subnormals below |x| ~ 0.06)
  hi-only  the `tc1` speed mode.
Result on the round-1 test weights (python tools/lo_scale_experiment.py): scaled 1.4e-7..5.5e-7, unscaled 1.9e-7..9.0e-7,
hi-only ~3e-4 -- the single-accumulator scheme keeps fp32-grade layers (ROUND2_NOTES.md)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lt_b200  # noqa: E402
from lt_b200 import testing  # noqa: E402

torch.manual_seed(0)
cfg = testing.make_config(num_layers=50, volume_size=32)
m = lt_b200.VolumetricTriangulationNet(cfg, device="cpu", backend="torch")
testing.randomize_weights(m, seed=0, calib_size=128, calib_views=1)
m = m.eval()
images, batch = testing.make_batch(1, 2, image_size=128, seed=0)
caps = []


def hook(mod, inp, out):
    if mod.weight.shape[1] >= 16:
        caps.append((mod, inp[0].detach()))


hooks = [mod.register_forward_hook(hook) for mod in m.modules() if isinstance(mod, (torch.nn.Conv2d, torch.nn.Conv3d))]
with torch.no_grad():
    m(images, None, batch)
for h in hooks:
    h.remove()


def split(x, scale):
    hi = x.half().float()
    return hi, ((x - hi) * scale).half().float()


def conv(mod, x, w):
    return (F.conv2d if w.dim() == 4 else F.conv3d)(x, w, None, mod.stride, mod.padding)


print("weight shape, median |x|, median |w|: rel err scaled / unscaled / hi-only")
for mod, x in caps[::max(1, len(caps) // 14)]:
    w = mod.weight.detach()
    y = conv(mod, x.double(), w.double())
    den = max(float(y.abs().max()), float(y.std()))
    res = []
    for scale in (2048.0, 1.0):
        xh, xl = split(x, scale)
        wh, wl = split(w, scale)
        yy = conv(mod, xh, wh).double() + (conv(mod, xh, wl).double() + conv(mod, xl, wh).double()) / scale
        res.append(float((yy - y).abs().max()) / den)
    res.append(float((conv(mod, split(x, 1.0)[0], split(w, 1.0)[0]).double() - y).abs().max()) / den)
    nz = x[x != 0].abs()
    print("%-26s %.2e %.2e: %.2e / %.2e / %.2e" % (tuple(w.shape), float(nz.median()), float(w.abs().median()), *res))
