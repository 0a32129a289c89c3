#!/usr/bin/env python
"""Per-kernel timings (CUDA events, warm-up, median of N) for the hot kernels at bench geometry.
Test tooling: prints one line per kernel with achieved TFLOP/s or GB/s.  Usage:
    python tools/bench_kernels.py [analysis] [gemm] [blstm] [--batch B]"""
import argparse
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops  # noqa: E402


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def bench_analysis(B):
    Bt, L, W, N = 3 * B, 64000, 1024, 256
    x = torch.randn(Bt, L, device="cuda") * 0.05
    filt = torch.randn(W, N, device="cuda") / 32
    flops = 2.0 * L * W * N * Bt
    for name, prec in (("bf16 tcgen05", ops.AMSS_PREC_BF16), ("fp32 simt", ops.AMSS_PREC_FP32)):
        if prec == ops.AMSS_PREC_FP32 and B > 16:
            continue
        med, best = timeit(lambda: ops.filterbank_analysis(x, filt, 256, 256, ops.AMSS_POOL_MAX, prec))
        print(f"analysis {name:14s} Bt={Bt}: {med:8.3f} ms median ({best:.3f} best)  {flops / med / 1e9:8.1f} TFLOP/s")
    src = torch.randn(B, 2, L, device="cuda") * 0.05
    xm = torch.cat([src[:, 0] + src[:, 1], src.reshape(2 * B, L)], 0).contiguous()
    med, best = timeit(lambda: ops.filterbank_analysis_mix(xm, filt, B, 2, 256, 256, ops.AMSS_PREC_BF16))
    print(f"analysis mix (linear mixture rows) B={B}: {med:8.3f} ms median ({best:.3f} best)  "
          f"{flops / med / 1e9:8.1f} TFLOP/s algorithmic, {flops * 2 / 3 / med / 1e9:8.1f} executed")


def bench_gemm(B):
    T = 250
    M = T * B
    shapes = [("blstm in-proj L1", M, 1200, 256, 0, 0), ("blstm in-proj L2", M, 1200, 600, 0, 0),
              ("head", M, 10240, 600, 0, 0), ("head dW", 600, 10240, M, 1, 0), ("head dH", M, 600, 10240, 0, 1),
              ("blstm dWx", 600, 1200, M, 1, 0), ("blstm dx", M, 600, 1200, 0, 1)]
    for name, m, n, k, ta, tb in shapes:
        A = torch.randn((k, m) if ta else (m, k), device="cuda")
        Bm = torch.randn((n, k) if tb else (k, n), device="cuda")
        out = torch.empty(m, n, device="cuda")
        for pname, prec in (("bf16", ops.AMSS_PREC_BF16), ("fp32", ops.AMSS_PREC_FP32)):
            try:
                med, best = timeit(lambda: ops.gemm(A, Bm, None, bool(ta), bool(tb), out=out, precision=prec), reps=5)
            except Exception as ex:  # noqa: BLE001
                print(f"gemm {name:18s} {pname}: {ex}")
                continue
            print(f"gemm {name:18s} {pname} M={m} N={n} K={k}: {med:8.3f} ms  {2.0 * m * n * k / med / 1e9:8.1f} TFLOP/s")


def bench_blstm(B):
    T = 250
    for I, H in ((256, 300), (600, 300)):
        x = torch.randn(T, B, I, device="cuda") * 0.1
        kf = torch.randn(I + H, 4 * H, device="cuda") * 0.05
        kb = torch.randn(I + H, 4 * H, device="cuda") * 0.05
        bf = torch.zeros(4 * H, device="cuda")
        bb = torch.zeros(4 * H, device="cuda")
        for pname, prec in (("bf16", ops.AMSS_PREC_BF16), ("fp32", ops.AMSS_PREC_FP32)):
            try:
                med, _ = timeit(lambda: ops.blstm_fwd(x, kf, bf, kb, bb, precision=prec), reps=5)
                y, saved = ops.blstm_fwd(x, kf, bf, kb, bb, precision=prec)
                dy = torch.randn_like(y)
                medb, _ = timeit(lambda: ops.blstm_bwd(x, kf, kb, y, dy, saved, precision=prec), reps=5)
            except Exception as ex:  # noqa: BLE001
                print(f"blstm I={I} H={H} {pname}: {ex}")
                continue
            print(f"blstm I={I} H={H} B={B} {pname}: fwd {med:8.3f} ms ({med / T * 1e3:.1f} us/step)  bwd {medb:8.3f} ms")


def bench_dpcl(B):
    TF, E = 64000, 40
    z = torch.randn(B, TF, E, device="cuda")
    lab = torch.randint(0, 2, (B, TF), device="cuda", dtype=torch.uint8)
    V, inv = ops.l2norm_fwd(z, E)
    one = torch.ones(1, device="cuda")
    gb = B * TF * E * 4 / 1e9
    loss, ws = ops.dpcl_loss_fwd(V, lab, 2)
    dV = ops.dpcl_loss_bwd(V, lab, 2, one, ws)
    for name, fn, passes in (("l2norm_fwd", lambda: ops.l2norm_fwd(z, E), 2), ("dpcl_loss_fwd", lambda: ops.dpcl_loss_fwd(V, lab, 2), 1),
                             ("dpcl_loss_fwd tc", lambda: ops.dpcl_loss_fwd(V, lab, 2, ops.AMSS_PREC_BF16), 1),
                             ("dpcl_loss_bwd", lambda: ops.dpcl_loss_bwd(V, lab, 2, one, ws), 2),
                             ("l2norm_bwd", lambda: ops.l2norm_bwd(V, inv, dV, E), 3),
                             ("dpcl_loss_bwd_normalized", lambda: ops.dpcl_loss_bwd_normalized(V, lab, 2, one, ws, inv), 2),
                             ("dpcl_bwd_normalized tc", lambda: ops.dpcl_loss_bwd_normalized(V, lab, 2, one, ws, inv, ops.AMSS_PREC_BF16), 2)):
        med, _ = timeit(fn, reps=5)
        print(f"{name:26s} B={B}: {med:8.3f} ms  ({passes * gb / med * 1e3:7.0f} GB/s algorithmic)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["analysis", "gemm", "blstm"])
    ap.add_argument("--batch", type=int, default=16)
    a = ap.parse_args()
    print(torch.cuda.get_device_name(0))
    if "analysis" in a.what:
        bench_analysis(a.batch)
    if "gemm" in a.what:
        bench_gemm(a.batch)
    if "blstm" in a.what:
        bench_blstm(a.batch)
    if "dpcl" in a.what:
        bench_dpcl(a.batch)
