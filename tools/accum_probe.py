#!/usr/bin/env python
"""How does tcgen05.mma kind::f16 round its fp32 accumulation?  D = A * B^T (plain fp16 operands, selftest path of conv_tc.cu) against
the exact float64 product, for all-positive operands (truncation shows as a negative mean error that grows with K) and for
zero-mean operands.  Errors in units of 2^-24 relative (fp32 half-ulp at the result magnitude)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lt_b200 import capi
torch.manual_seed(0)
M, N = 128, 64
for K in (64, 256, 1024, 4096, 16384):
    for kind in ("positive", "zero-mean"):
        a = torch.rand(M, K, device="cuda") + (0.5 if kind == "positive" else -0.5)
        b = torch.rand(N, K, device="cuda") + (0.5 if kind == "positive" else -0.5)
        a, b = a.half(), b.half()
        d = torch.zeros(M * N + 2 * N, dtype=torch.float32, device="cuda")
        capi.tc_gemm_selftest(a, b, d, M, N, K)
        torch.cuda.synchronize()
        got = d[:M * N].view(M, N).double()
        want = a.double() @ b.double().t()
        f32 = (a.float() @ b.float().t()).double()     # cuBLAS fp32 (FFMA / its own order) for scale
        scale = want.abs().mean() if kind == "positive" else (a.double().abs() @ b.double().abs().t()).mean() / K ** 0.5
        err = (got - want) / scale / 2.0 ** -24
        err32 = (f32 - want) / scale / 2.0 ** -24
        print("K %6d %-9s tcgen05: mean %+8.2f rms %7.2f | torch fp32 matmul: mean %+8.2f rms %7.2f   (units of 2^-24 x result scale)"
              % (K, kind, float(err.mean()), float(err.pow(2).mean().sqrt()), float(err32.mean()), float(err32.pow(2).mean().sqrt())))

# How much of the truncation error is a pure shrinkage of the result (err = -beta * result)?  If most of it is, a per-layer factor
# (1 + beta(K)) folded into the epilogue scale would remove it.  Data model of a conv layer: post-ReLU activations x zero-mean weights.
print("\nshrinkage model: err ~ -beta * result (relu(normal) activations x normal weights)")
for K in (256, 1024, 2304, 4608):
    a = torch.relu(torch.randn(M, K, device="cuda")).half()
    b = (torch.randn(N, K, device="cuda") * (2.0 / K) ** 0.5 * 700).half()     # Kaiming-sized weights after the power-of-two pre-scale
    d = torch.zeros(M * N + 2 * N, dtype=torch.float32, device="cuda")
    capi.tc_gemm_selftest(a, b, d, M, N, K)
    torch.cuda.synchronize()
    got = d[:M * N].view(M, N).double()
    want = a.double() @ b.double().t()
    err = got - want
    beta = -float((err * want).sum() / (want * want).sum())
    resid = err + beta * want
    scale = float(want.pow(2).mean().sqrt())
    print("K %5d  beta %.3e (= %.4f * K/16 * 2^-24)  rms err / rms result: raw %.3e, after (1 + beta) correction %.3e" %
          (K, beta, beta / (K / 16 * 2.0 ** -24), float(err.pow(2).mean().sqrt()) / scale, float(resid.pow(2).mean().sqrt()) / scale))
