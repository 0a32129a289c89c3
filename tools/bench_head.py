#!/usr/bin/env python
"""Timing of the embedding-head GEMM variants (plain / fused l2-normalise epilogue) at bench geometry.  Diagnostics."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402

M, N, K = 32000, 10240, 600
x = torch.randn(M, K, device="cuda")
W = torch.randn(K, N, device="cuda") * 0.05
b = torch.zeros(N, device="cuda")
xb, Wb = ops.convert_bf16(x), ops.convert_bf16(W)
out = torch.empty(M, N, device="cuda")
for name, fn in (("plain", lambda: ops.gemm_bf16(xb, False, Wb, True, M, N, K, bias=b, out=out)),
                 ("plain, no bias", lambda: ops.gemm_bf16(xb, False, Wb, True, M, N, K, out=out)),
                 ("norm E=40", lambda: ops.gemm_bf16(xb, False, Wb, True, M, N, K, bias=b, out=out, norm_E=40)),
                 ("norm E=16", lambda: ops.gemm_bf16(xb, False, Wb, True, M, N, K, bias=b, out=out, norm_E=16)),
                 ("norm E=8", lambda: ops.gemm_bf16(xb, False, Wb, True, M, N, K, bias=b, out=out, norm_E=8)),
                 ("norm E=40 no bias", lambda: ops.gemm_bf16(xb, False, Wb, True, M, N, K, out=out, norm_E=40))):
    med, best = timeit(fn, reps=5)
    print(f"head {name:18s}: {med:8.3f} ms ({2.0 * M * N * K / med / 1e9:7.1f} TFLOP/s)")
