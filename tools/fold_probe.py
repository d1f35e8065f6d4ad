"""Time one kw-folded conv (3^3 32->32 and 7^3 32->16 at 64^3, B=8) under the lt_options.fold_debug settings (LT_OPT_FOLD_DEBUG)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for dbg in ("0", "1", "2", "3", "0p"):   # "0p": production path with the in-kernel wait counters (fold_debug = 16)
        env = dict(os.environ, LT_OPT_FOLD_DEBUG=dbg[0])
        if dbg.endswith("p"):
            env["LT_OPT_FOLD_DEBUG"] = "16"
        r = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True)
        print("fold_debug=" + dbg, r.stdout.strip())
        print("\n".join(sorted(set(l for l in r.stderr.strip().splitlines() if "fold prof" in l))[:8]))
    sys.exit(0)
import torch
from lt_b200 import capi
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_ops import _engine, _bn_for, act_from_nchw
e = _engine("tc")
out = []
for cin, cout, k in ((32, 32, 3), (32, 16, 7)):
    conv = torch.nn.Conv3d(cin, cout, k, 1, k // 2).eval().cuda()
    bn = _bn_for(conv, 1).cuda()
    pk = e._pack_conv(conv, bn, cin_pad=32)
    x = act_from_nchw(torch.randn(8, cin, 64, 64, 64, device="cuda"), capi.FMT_S32, pad_c=32)
    for _ in range(2):
        e._conv(x, pk, relu=True)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        e._conv(x, pk, relu=True)
    t1.record()
    torch.cuda.synchronize()
    out.append("k%d: %.3f ms" % (k, t0.elapsed_time(t1) / 5))
print("  ".join(out))
