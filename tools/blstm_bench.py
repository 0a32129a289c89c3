#!/usr/bin/env python
"""Timing of the tcgen05 BLSTM layer (amss_blstm_fwd / amss_blstm_bwd, bench geometry T = 250, I = 600, H = 300) for both
sub-batch variants of the recurrence (AMSS_BLSTM_NB = 16: two CTAs per SM; 32: one), plus the co-resident cluster counts."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops, _lib  # noqa: E402

out4 = (ctypes.c_int * 4)()
_lib.call("amss_debug_blstm_clusters", 300, ctypes.addressof(out4))
print("co-resident clusters H=300 {fwd16, fwd32, bwd16, bwd32}:", list(out4))
out4b = (ctypes.c_int * 4)()
_lib.call("amss_debug_blstm_clusters", 150, ctypes.addressof(out4b))
print("co-resident clusters H=150:", list(out4b))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


T, I, H = 250, 600, 300
for B in (32, 64, 128, 256):
    x = torch.randn(T, B, I, device="cuda") * 0.1
    kf = torch.randn(I + H, 4 * H, device="cuda") * 0.05
    kb = torch.randn(I + H, 4 * H, device="cuda") * 0.05
    bf = torch.zeros(4 * H, device="cuda")
    dy = torch.randn(T, B, 2 * H, device="cuda")
    for nb in ("16", "32", ""):
        if nb:
            os.environ["AMSS_BLSTM_NB"] = nb
        else:
            os.environ.pop("AMSS_BLSTM_NB", None)
        y, saved = ops.blstm_fwd(x, kf, bf, kb, bf, precision=ops.AMSS_PREC_BF16)
        f = timeit(lambda: ops.blstm_fwd(x, kf, bf, kb, bf, precision=ops.AMSS_PREC_BF16))
        b = timeit(lambda: ops.blstm_bwd(x, kf, kb, y, dy, saved, precision=ops.AMSS_PREC_BF16))
        print(f"B={B:4d} NB={nb or 'auto':>4s}: blstm_fwd {f:7.3f} ms  blstm_bwd {b:7.3f} ms")
