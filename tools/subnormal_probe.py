#!/usr/bin/env python
"""Does tcgen05.mma kind::f16 honour fp16 SUBNORMAL operands?  D = A * B^T with A entries in the subnormal range (2^-20) and B = 2^10:
exact result K * 2^-10.  Flush-to-zero would return 0."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lt_b200 import capi
M, N, K = 128, 64, 64
for name, aval, bval in (("normal", 2.0 ** -10, 1.0), ("A subnormal 2^-20", 2.0 ** -20, 2.0 ** 10), ("B subnormal 2^-20", 2.0 ** 10, 2.0 ** -20),
                         ("A = 2^-24 (smallest subnormal)", 2.0 ** -24, 2.0 ** 10), ("mixed: 1 + subnormal partner", None, None)):
    if aval is None:
        a = torch.ones(M, K, dtype=torch.float16, device="cuda")
        a[:, 1::2] = 2.0 ** -18          # every second K element subnormal
        b = torch.ones(N, K, dtype=torch.float16, device="cuda")
        want = K / 2 * 1.0 + K / 2 * 2.0 ** -18
    else:
        a = torch.full((M, K), aval, dtype=torch.float16, device="cuda")
        b = torch.full((N, K), bval, dtype=torch.float16, device="cuda")
        want = K * aval * bval
    d = torch.zeros(M * N + 2 * N, dtype=torch.float32, device="cuda")
    capi.tc_gemm_selftest(a, b, d, M, N, K)
    torch.cuda.synchronize()
    got = d[:M * N]
    print("%-34s want %.9g got min %.9g max %.9g" % (name, want, float(got.min()), float(got.max())))
