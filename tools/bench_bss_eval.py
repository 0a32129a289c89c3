#!/usr/bin/env python
"""BSS-eval throughput of the batched float64 device implementation.
    python tools/bench_bss_eval.py [--batch 16] [--len 64000]
(The per-mixture numpy loop of the reference's vendored mir_eval code runs at ~6 mixtures/s on the same box at
L = 64000; it lives in oracle/, which only tests / smoke / the bench's CPU leg may import.)"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import bss_eval as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--len", type=int, default=64000)
a = ap.parse_args()
rng = np.random.RandomState(0)
ref = rng.randn(a.batch, 2, a.len)
est = ref[:, ::-1] * 0.9 + 0.1 * ref + 0.05 * rng.randn(a.batch, 2, a.len)
r, e = torch.tensor(ref, device="cuda"), torch.tensor(est.copy(), device="cuda")
G.bss_eval_sources(r, e)
torch.cuda.synchronize()
t0 = time.time()
out = G.bss_eval_sources(r, e)
torch.cuda.synchronize()
tg = time.time() - t0
print(f"device: {a.batch / tg:.1f} mixtures/s ({tg * 1e3:.1f} ms for {a.batch});  sdr of mixture 0: {out[0][0].cpu().numpy()}")
