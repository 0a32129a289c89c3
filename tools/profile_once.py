#!/usr/bin/env python
"""One launch of each hot kernel at bench geometry, for `ncu --set full` captures (diagnostics)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
from amss_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T, L = 250, 64000
P = ops.AMSS_PREC_BF16
torch.manual_seed(0)
x = torch.randn(3 * B, L, device="cuda") * 0.05
filt = torch.randn(1024, 256, device="cuda") / 32
ops.filterbank_analysis(x, filt, 256, 256, ops.AMSS_POOL_MAX, P)
src = x[B:].reshape(B, 2, L)
xm = torch.cat([src[:, 0] + src[:, 1], x[B:]], 0).contiguous()
ops.filterbank_analysis_mix(xm, filt, B, 2, 256, 256, P)                                # linear-mixture fast path
M = T * B
h = torch.randn(M, 600, device="cuda")
W = torch.randn(600, 10240, device="cuda") * 0.05
bias = torch.zeros(10240, device="cuda")
hb, Wb = ops.convert_bf16(h), ops.convert_bf16(W)
V, inv = ops.gemm_bf16(hb, False, Wb, True, M, 10240, 600, bias=bias, norm_E=40)      # head fwd + fused l2-normalise
dzb = ops.convert_bf16(torch.randn(M, 10240, device="cuda"))
ops.gemm_bf16(hb, True, dzb, True, 600, 10240, M)                                     # head dW (split-K, add-reduce stores)
ops.gemm_bf16(dzb, False, Wb, False, M, 600, 10240)                                   # head dH
ops.gemm_bf16(hb, False, ops.convert_bf16(torch.randn(600, 1200, device="cuda")), True, M, 1200, 600)   # BLSTM in-proj
xt = torch.randn(T, B, 600, device="cuda") * 0.1
kf = torch.randn(900, 1200, device="cuda") * 0.05
bf = torch.zeros(1200, device="cuda")
y, saved = ops.blstm_fwd(xt, kf, bf, kf, bf, precision=P)
ops.blstm_bwd(xt, kf, kf, y, torch.randn_like(y), saved, precision=P)
lab = torch.randint(0, 2, (B, 64000), device="cuda", dtype=torch.uint8)
Vn, invn = ops.l2norm_fwd(torch.randn(B, 64000, 40, device="cuda"), 40)
loss, ws = ops.dpcl_loss_fwd(Vn, lab, 2, P)                                           # tcgen05 Gram
ops.dpcl_loss_bwd_normalized_bf16(Vn, lab, 2, torch.ones(1, device="cuda"), ws, invn)  # fused DPCL + l2norm backward, bf16 dz
torch.cuda.synchronize()
print("done")
