#!/usr/bin/env python
"""Forward STFT / label kernels at the bench geometry (L = 64000, frame 512, hop 256): per-frame kernels
(AMSS_STFT_PER_FRAME=1) vs several frames per CTA with two frames per complex FFT."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops  # noqa: E402


def timed(fn, n=20):
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


for B, S in ((64, 3), (32, 2), (4, 2)):
    x = torch.randn(B, 64000, device="cuda")
    xs = torch.randn(B, S, 64000, device="cuda")
    for mode in ("1", None):
        if mode:
            os.environ["AMSS_STFT_PER_FRAME"] = mode
        else:
            os.environ.pop("AMSS_STFT_PER_FRAME", None)
        a = timed(lambda: ops.stft(x, 512, 256))
        b = timed(lambda: ops.stft_labels(xs, 512, 256, True))
        print(f"B={B} S={S} {'per-frame' if mode else 'runs     '}: stft {a * 1e3:7.1f} us   labels+mag {b * 1e3:7.1f} us")
    os.environ["AMSS_STFT_PER_FRAME"] = "1"
    s0, m0 = ops.stft(x, 512, 256)
    l0, g0 = ops.stft_labels(xs, 512, 256, True)
    os.environ.pop("AMSS_STFT_PER_FRAME")
    s1, m1 = ops.stft(x, 512, 256)
    l1, g1 = ops.stft_labels(xs, 512, 256, True)
    print(f"   max |spec diff| {float((s0 - s1).abs().max()):.2e} (|spec| max {float(s0.abs().max()):.1f}), labels equal "
          f"{float((l0 == l1).float().mean()) * 100:.4f} %, mag diff {float((g0 - g1).abs().max()):.2e}")
