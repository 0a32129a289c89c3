#!/usr/bin/env python
"""Schedule trace of the tcgen05 BLSTM recurrence kernels (diagnostics, amss_debug_blstm_sched): which SM every CTA ran on and
when each cluster started / ended, i.e. how many clusters were really co-resident and how many waves a launch took."""
import os
import sys
from collections import Counter

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops, _lib  # noqa: E402

T, I, H = 250, 600, 300
NC = (H + 31) // 32
for B in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    x = torch.randn(T, B, I, device="cuda") * 0.1
    kf = torch.randn(I + H, 4 * H, device="cuda") * 0.05
    kb = torch.randn(I + H, 4 * H, device="cuda") * 0.05
    bf = torch.zeros(4 * H, device="cuda")
    for _ in range(2):
        y, saved = ops.blstm_fwd(x, kf, bf, kb, bf, precision=ops.AMSS_PREC_BF16)
        ops.blstm_bwd(x, kf, kb, y, torch.randn_like(y), saved, precision=ops.AMSS_PREC_BF16)
    buf = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
    _lib.call("amss_debug_blstm_sched", buf.data_ptr())
    y, saved = ops.blstm_fwd(x, kf, bf, kb, bf, precision=ops.AMSS_PREC_BF16)
    ops.blstm_bwd(x, kf, kb, y, torch.randn_like(y), saved, precision=ops.AMSS_PREC_BF16)
    torch.cuda.synchronize()
    _lib.call("amss_debug_blstm_sched", 0)
    h = buf.cpu().view(2, 4096, 4)
    for name, tr in (("fwd", h[0]), ("bwd", h[1])):
        n = int((tr[:, 1] > 0).sum())
        tr = tr[:n]
        t0 = int(tr[:, 1].min())
        per_sm = Counter(int(v) for v in tr[:, 0])
        print(f"B={B} {name}: {n} CTAs = {n // NC} clusters on {len(per_sm)} SMs (max {max(per_sm.values())} CTAs on one SM); "
              f"kernel {(int(tr[:, 2].max()) - t0) / 1e3:.1f} us")
        for c in range(n // NC):
            rows = tr[c * NC:(c + 1) * NC]
            sms = sorted(int(v) for v in rows[:, 0])
            print(f"   cluster {c:2d}: start +{(int(rows[:, 1].min()) - t0) / 1e3:7.1f} us  prologue done +{(int(rows[:, 3].max()) - t0) / 1e3:6.1f} us"
                  f"  end +{(int(rows[:, 2].max()) - t0) / 1e3:7.1f} us  SMs {sms}")
