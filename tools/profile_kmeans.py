#!/usr/bin/env python
"""One k-means fit at the config-5 geometry (TF = 63993, E = 40, K = 3, 10 tries x 10 steps) for `ncu` captures."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
TF, E, tries, iters = 63993, 40, 10, 10
torch.manual_seed(0)
X = torch.randn(B, TF, E, device="cuda")
idx = torch.as_tensor(np.random.RandomState(0).randint(0, TF, size=(B * tries, K)).astype(np.int32)).cuda()
for _ in range(2):
    ops.kmeans_fit(X, idx, K, tries, iters)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.kmeans_fit(X, idx, K, tries, iters)
e1.record()
torch.cuda.synchronize()
print(f"kmeans_fit B={B} K={K}: {e0.elapsed_time(e1):.3f} ms")
