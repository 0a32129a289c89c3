#!/usr/bin/env python
"""Timing of the sparse filterbank kernels at the config-4 geometry (adapt pre-training: B mixtures, S = 2, L = 64000, W = 1024,
N = 256 filters, max-pool 256): analysis backward (filter gradient through the arg-max), synthesis forward (sparse overlap-add),
synthesis backward (dvals and filter gradient).  Prints ms and the fp32 FMA rate (every atom = W multiply-adds)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops  # noqa: E402


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S, L, W, N, pool = 2, 64000, 1024, 256, 256
torch.manual_seed(0)
src = torch.randn(B, S, L, device="cuda") * 0.05
x = torch.cat([src.sum(1), src.reshape(B * S, L)], 0).contiguous()                      # [B(S+1), L]
filt = torch.randn(W, N, device="cuda") / 32
y, am = ops.filterbank_analysis(x, filt, pool, pool, ops.AMSS_POOL_MAX, ops.AMSS_PREC_BF16)
Tp = y.shape[1]
dy = torch.randn_like(y)
fma = lambda rows: rows * Tp * N * W                                                     # noqa: E731
t = timeit(lambda: ops.filterbank_analysis_bwd(x, dy, am, W))
print(f"analysis backward  ({x.shape[0]} rows): {t:7.3f} ms  {fma(x.shape[0]) / t / 1e9:7.2f} TFMA/s")
vals = y[B:].contiguous()                                                               # [B*S, Tp, N] source rows
am_mix = am[:B].contiguous()
t = timeit(lambda: ops.filterbank_synthesis(vals, am_mix, filt, B, S, L, pool, pool))
print(f"synthesis forward  ({B * S} rows): {t:7.3f} ms  {fma(B * S) / t / 1e9:7.2f} TFMA/s")
dout = torch.randn(B * S, L, device="cuda")
t = timeit(lambda: ops.filterbank_synthesis_bwd(dout, vals, am_mix, filt, B, S, need_dvals=True, need_dfilt=False))
print(f"synthesis bwd dvals ({B * S} rows): {t:7.3f} ms  {fma(B * S) / t / 1e9:7.2f} TFMA/s")
t = timeit(lambda: ops.filterbank_synthesis_bwd(dout, vals, am_mix, filt, B, S, need_dvals=False, need_dfilt=True))
print(f"synthesis bwd dfilt ({B * S} rows): {t:7.3f} ms  {fma(B * S) / t / 1e9:7.2f} TFMA/s")
