#!/usr/bin/env python
"""Time single conv layers of the config #2 step in isolation (B200): residual on/off, L2 warm / flushed, kernel choice.

    python tools/conv_probe.py                       # default case list (the layers that dominate the step)
    python tools/conv_probe.py 256,1024,1,1,24,32    # cin,cout,k,stride,side,N  [more cases ...]
Per case: 3 warm-up + 20 timed launches (CUDA events); "flushed" writes a 256 MB buffer between launches, "warm" does not
(input / residual / weights then sit in L2 as they do behind the producing layer inside the step).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from lt_b200 import capi  # noqa: E402
from lt_b200.engine import Act  # noqa: E402
from test_gpu_ops import _engine, _bn_for, DEV  # noqa: E402

CASES = [(256, 1024, 1, 1, 24, 32), (1024, 256, 1, 1, 24, 32), (256, 256, 3, 1, 24, 32), (128, 512, 1, 1, 48, 32), (64, 256, 1, 1, 96, 32),
         (64, 64, 3, 1, 96, 32)]
if len(sys.argv) > 1:
    CASES = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=DEV)


def timed(fn, do_flush, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


for cin, cout, k, stride, side, N in CASES:
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2, bias=False).eval().to(DEV)
    bn = _bn_for(conv, 1).to(DEV)
    out_side = (side + 2 * (k // 2) - k) // stride + 1
    x = Act(N, 1, side, side, cin, capi.FMT_S32, DEV)
    x.data.normal_()
    res = Act(N, 1, out_side, out_side, cout, capi.FMT_S32, DEV)
    res.data.normal_()
    out = Act(N, 1, out_side, out_side, cout, capi.FMT_S32, DEV)
    flops = 2.0 * N * out_side * out_side * cin * cout * k * k
    mb_min = (x.data.numel() + out.data.numel()) * 2 / 1e6
    print("cin %4d cout %4d k%d s%d %3dx%-3d N%-3d  [in+out %.0f MB, %.1f GFLOP]" % (cin, cout, k, stride, side, side, N, mb_min, flops / 1e9), flush=True)
    variants = [("pair direct-store ", True, dict(pair_direct_out=1)), ("pair staged+TMA   ", True, dict(pair_direct_out=0)),
                ("pair Nt=128        ", True, dict(pair_direct_out=1, pair_nt=128)), ("pair ring depth 2  ", True, dict(pair_direct_out=1, pair_stages=2)),
                ("one-CTA kernel    ", False, dict())]
    for name, use_pair, o in variants:
        capi.set_options(pair_direct_out=1, pair_nt=0, pair_stages=0, pair_prof=0)
        capi.set_options(**o)
        e = _engine("tc")
        e.use_pair = use_pair
        pk = e._pack_conv(conv, bn)
        line = "   %s:" % name
        for with_res in (False, True):
            r, mode = (res, capi.RES_BEFORE_RELU) if with_res else (None, capi.RES_NONE)
            fn = lambda: e._conv(x, pk, relu=True, residual=r, res_mode=mode, out=out)   # noqa: E731
            for do_flush in (False, True):
                us = timed(fn, do_flush)
                line += "  %s%s %6.1f us (%4.0f TF/s)" % ("+res" if with_res else "    ", " flushed" if do_flush else " warm   ", us, flops / us / 1e6)
        print(line, flush=True)
        if use_pair and os.environ.get("CONV_PROBE_PROF", "1") == "1":
            capi.set_options(pair_prof=1)
            for with_res in (False, True):
                r, mode = (res, capi.RES_BEFORE_RELU) if with_res else (None, capi.RES_NONE)
                flush.zero_()
                torch.cuda.synchronize()
                sys.stderr.write("      %s %s: " % (name, "+res" if with_res else "no res"))
                sys.stderr.flush()
                e._conv(x, pk, relu=True, residual=r, res_mode=mode, out=out)
                torch.cuda.synchronize()
            capi.set_options(pair_prof=0)
capi.set_options(pair_direct_out=1, pair_nt=0, pair_stages=0, pair_prof=0)
