#!/usr/bin/env python
"""Tile profile of the tcgen05 k-means update pass (diagnostics): clock deltas between the stamps of
amss_debug_kmeans_profile for tiles 8..11 of CTA 0 (loader thread 0, MMA issuer, first epilogue thread)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
TF, E, tries, iters = 63993, 40, 10, 10
torch.manual_seed(0)
X = torch.randn(B, TF, E, device="cuda")
idx = torch.as_tensor(np.random.RandomState(0).randint(0, TF, size=(B * tries, K)).astype(np.int32)).cuda()
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
ops.kmeans_fit(X, idx, K, tries, iters)
_lib.call("amss_debug_kmeans_profile", buf.data_ptr())
ops.kmeans_fit(X, idx, K, tries, iters)
torch.cuda.synchronize()
_lib.call("amss_debug_kmeans_profile", 0)
p = buf.cpu().view(-1)[:96].view(3, 4, 8)
names = [["top", "raw_full", "row read, |x|^2", "normalised", "x3_empty+d1_empty", "split+stores issued",
          "tmem_st_wait", "arrive"],
         ["top", "x3_full", "d1_empty", "P1 issued", "oh_full(i-1)", "P2(i-1) issued", "-", "-"],
         ["top", "x3_full", "d1_full", "tmem ld", "distances", "oh_empty", "one-hot stored", "-"]]
for role, rn in enumerate(("loader", "mma", "epilogue")):
    print(f"--- {rn}")
    for t in range(3):
        row = p[role][t]
        base = int(row[0])
        # the loader groups take alternate tiles: the period of a loader thread spans two tiles
        nxt = int(p[role][t + 1][0]) if role else (int(p[role][t + 2][0]) if t < 2 else base)
        print(f" tile {8 + t}: {'two-tile period' if role == 0 else 'total'} {nxt - base} clk ; " +
              " ".join(f"[{names[role][k]}]+{int(row[k]) - base}" for k in range(1, 8) if names[role][k] != "-" and int(row[k]) != 0))
