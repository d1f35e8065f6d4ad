"""Time the unprojection and the soft-argmax in isolation (config #2 shapes, B = 8) under kernel-selection knobs.

    python tools/post_probe.py            # runs every variant in a subprocess and prints one line each
Each measurement: 3 warm-up + 10 timed launches, CUDA events, a 256 MB L2-flush write between launches.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = [
    ("unproject", {"LT_OPT_UNPROJECT_V2": "0"}), ("unproject", {"LT_OPT_UNPROJECT_CPL": "4"}), ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_LB": "5"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_BRICK": "0"}), ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_BRICK": "4"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_BRICK": "8", "LT_OPT_UNPROJECT_BRICK_ORDER": "0"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_BRICK": "8", "LT_OPT_UNPROJECT_BRICK_ORDER": "1"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_BRICK": "8", "LT_OPT_UNPROJECT_BRICK_ORDER": "2"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_BRICK": "4", "LT_OPT_UNPROJECT_BRICK_ORDER": "2"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_BRICK": "16", "LT_OPT_UNPROJECT_BRICK_ORDER": "2"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "4", "LT_OPT_UNPROJECT_BRICK": "16", "LT_OPT_UNPROJECT_BRICK_ORDER": "1"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "8", "LT_OPT_UNPROJECT_LB": "3", "LT_OPT_UNPROJECT_BRICK": "8", "LT_OPT_UNPROJECT_BRICK_ORDER": "1"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "8", "LT_OPT_UNPROJECT_LB": "3", "LT_OPT_UNPROJECT_BRICK": "8", "LT_OPT_UNPROJECT_BRICK_ORDER": "2"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "8"}), ("unproject", {"LT_OPT_UNPROJECT_CPL": "8", "LT_OPT_UNPROJECT_LB": "3"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "8", "LT_OPT_UNPROJECT_LB": "3", "LT_OPT_UNPROJECT_BRICK": "4"}),
    ("unproject", {"LT_OPT_UNPROJECT_CPL": "8", "LT_OPT_UNPROJECT_LB": "3", "LT_OPT_UNPROJECT_BRICK": "0"}),
    ("softargmax20", {"LT_OPT_SOFTARGMAX_STREAM": "0"}), ("softargmax20", {"LT_OPT_SOFTARGMAX_STREAM": "1"}),
    ("softargmax32", {"LT_OPT_SOFTARGMAX_STREAM": "0"}), ("softargmax32", {"LT_OPT_SOFTARGMAX_STREAM": "1"}),
]

if len(sys.argv) == 1:
    for what, env in VARIANTS:
        r = subprocess.run([sys.executable, __file__, what], env=dict(os.environ, **env), capture_output=True, text=True)
        print(what, env, r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from lt_b200 import capi, testing  # noqa: E402

dev = "cuda:0"
B, V, n, h, J = 8, 4, 64, 96, 17
nvox = n ** 3
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timed(fn, nbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 10
    return "%.4f ms  %.0f GB/s algorithmic" % (ms, nbytes / ms / 1e6)


what = sys.argv[1]
if what == "unproject":
    import lt_b200
    cfg = testing.make_config(num_layers=50, volume_size=n)
    model = lt_b200.VolumetricTriangulationNet(cfg, device=dev, backend="native", use_cuda_graph=False)
    _, batch = testing.make_batch(B, V, image_size=384, seed=0)
    proj, base, position, step, rots, _ = model._host_geometry(batch, B, (384, 384), (h, h))
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
    coord = torch.empty((B, n, n, n, 3), device=dev)
    capi.coord_volume(up(position), up(base), up(step), up(rots.reshape(B, 9)), coord, False)
    feats = torch.randn(B, V, h, h, 32, device=dev)
    out = torch.empty((B, nvox, 64), dtype=torch.float16, device=dev)
    pj = up(proj)
    nbytes = B * (nvox * 32 * 4 + V * h * h * 32 * 4 + nvox * 12)
    res = timed(lambda: capi.unproject_aggregate(feats, pj, coord.view(B, nvox, 3), None, out, capi.FMT_S32, capi.AGG["softmax"]), nbytes)
    f = torch.empty((B, nvox, 32), device=dev)
    capi.s32_to_f32(out, f, B * nvox, 32)
    print(res, " checksum: sum %.6f  sumsq %.6f  sample %s" % (float(f.double().sum()), float((f.double() ** 2).sum()),
                                                               [round(float(x), 6) for x in f[3, 12345:12348, 7]]))
else:
    vs = int(what[-2:])
    logits = torch.randn(B, nvox, vs, device=dev) * 3
    coord = torch.randn(B, nvox, 3, device=dev) * 700
    vol = torch.empty((B, J, nvox), device=dev)
    kp = torch.empty((B, J, 3), device=dev)
    ws = torch.empty(capi.softargmax3d_workspace_bytes(B, J, nvox) // 4 + 1, dtype=torch.float32, device=dev)
    nbytes = B * (2 * J * nvox * 4 + nvox * 12)
    print(timed(lambda: capi.softargmax3d(logits, nvox * vs, vs, 1, coord, vol, kp, ws, B, J, nvox, 1.0, True), nbytes))
