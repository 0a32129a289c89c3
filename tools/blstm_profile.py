#!/usr/bin/env python
"""Step profile of the tcgen05 BLSTM forward recurrence (diagnostics): prints clock deltas between the
stamps of amss_debug_blstm_profile for steps 100..103 of CTA 0."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops, _lib  # noqa: E402

NAMES = ["0 step start", "1 mma done", "2 tmem ld done", "3 activation done", "4 named bar", "5 cell done",
         "6 staged+arrive", "7 writer: barrier passed", "8 writer: gates stored", "9 mma warp: h landed", "10 mma issued",
         "11 writer: c,y stored"]

for B in [int(a) for a in sys.argv[1:]] or [16, 64, 128]:
    T, I, H = 250, 600, 300
    x = torch.randn(T, B, I, device="cuda") * 0.1
    kf = torch.randn(I + H, 4 * H, device="cuda") * 0.05
    kb = torch.randn(I + H, 4 * H, device="cuda") * 0.05
    bf = torch.zeros(4 * H, device="cuda")
    buf = torch.zeros(128, dtype=torch.int64, device="cuda")
    _lib.call("amss_debug_blstm_profile", buf.data_ptr())
    for _ in range(3):
        y, saved = ops.blstm_fwd(x, kf, bf, kb, bf, precision=ops.AMSS_PREC_BF16)
        ops.blstm_bwd(x, kf, kb, y, torch.randn_like(y), saved, precision=ops.AMSS_PREC_BF16)
    torch.cuda.synchronize()
    _lib.call("amss_debug_blstm_profile", 0)
    p = buf.cpu().view(-1)[:48].view(4, 12)
    print(f"--- B={B}")
    for s in range(1, 3):
        row = p[s]
        base = int(row[0])
        print(f" step {100 + s}: total {int(p[s + 1][0]) - base} clk ; " +
              " ".join(f"[{k}]+{int(row[k]) - base}" for k in (1, 2, 3, 5, 6, 4)) + " | mma warp " +
              " ".join(f"[{k}]+{int(row[k]) - base}" for k in (9, 10)))
        w7 = int(row[7])   # writer stamps come from another SM sub-partition (own clock offset): differences only
        print(f"      writer: zx issued +{int(row[11]) - w7}, tiles read +{int(buf.cpu()[48 + s]) - w7}, stores issued +{int(row[8]) - w7}"
              f" ; barrier-to-barrier {int(p[s + 1][7]) - w7}")

    pb = buf.cpu().view(-1)[64:112].view(4, 12)
    names = {9: "mma issue", 10: "issued", 1: "factors ready", 2: "mma done", 11: "tmem read", 8: "staged", 3: "pushed", 4: "partials landed", 5: "reduced",
             6: "dz staged", 7: "barrier"}
    for s in range(1, 3):
        row = pb[s]
        base = int(row[0])
        print(f" bwd step {100 + s}: total {int(pb[s + 1][0]) - base} clk ; " +
              " ".join(f"{names[k]} +{int(row[k]) - base}" for k in (9, 10, 1, 2, 11, 8, 3, 4, 5, 6, 7)))
