"""End-to-end CPU emulation of the split-fp16 3-product convolutions (every Conv / ConvTranspose of the calibrated ResNet-50 +
V2V test model, 2 views 128x128, 32^3) with scaled (x = hi + lo/2048) and UNSCALED (x = hi + lo) low parts against the fp32
forward.  Round-1 result: scaled 0.0245 mm / features 4.3e-6 / volumes 3.5e-5; unscaled 0.0291 mm / 4.5e-6 / 3.9e-5, arg-max
voxels equal in both -- the single-accumulator scheme is safe end to end (ROUND2_NOTES.md)."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, types, numpy as np, torch.nn.functional as F
import lt_b200
from lt_b200 import testing
torch.manual_seed(0)
cfg = testing.make_config(num_layers=50, volume_size=32)
m = lt_b200.VolumetricTriangulationNet(cfg, device="cpu", backend="torch")
testing.randomize_weights(m, seed=0, calib_size=128, calib_views=1)
m = m.eval()
images, batch = testing.make_batch(1, 2, image_size=128, seed=0)
with torch.no_grad():
    ref = m(images, None, batch)
def split(x, scale):
    x = x.clamp(-65504, 65504)
    hi = x.half().float()
    return hi, ((x - hi) * scale).half().float()
def make_forward(mod, scale):
    def fwd(self, x):
        xh, xl = split(x, scale); wh, wl = split(self.weight, scale)
        kw = dict(stride=self.stride, padding=self.padding)
        if isinstance(self, torch.nn.ConvTranspose2d) or isinstance(self, torch.nn.ConvTranspose3d):
            f = F.conv_transpose2d if x.dim() == 4 else F.conv_transpose3d
        else:
            f = F.conv2d if x.dim() == 4 else F.conv3d
        y = f(xh, wh, None, **kw) + (f(xh, wl, None, **kw) + f(xl, wh, None, **kw)) / scale
        if self.bias is not None:
            y = y + self.bias.view(1, -1, *([1] * (y.dim() - 2)))
        return y
    return types.MethodType(fwd, mod)
for scale in (2048.0, 1.0):
    saved = []
    for mod in m.modules():
        if isinstance(mod, (torch.nn.Conv2d, torch.nn.Conv3d, torch.nn.ConvTranspose2d, torch.nn.ConvTranspose3d)):
            saved.append((mod, mod.forward)); mod.forward = make_forward(mod, scale)
    with torch.no_grad():
        out = m(images, None, batch)
    for mod, f in saved: mod.forward = f
    def rel(a, b): return float((a - b).abs().max() / max(float(b.abs().max()), float(b.std())))
    print("lo scale %6.0f: keypoints max err %.4f mm, features rel %.2e, volumes rel %.2e, argmax equal %s" % (
        scale, float((out[0] - ref[0]).abs().max()), rel(out[1], ref[1]), rel(out[2], ref[2]),
        bool(torch.equal(out[2].reshape(1, 17, -1).argmax(-1), ref[2].reshape(1, 17, -1).argmax(-1)))))
