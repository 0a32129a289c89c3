#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the CPU oracle (oracle/).

The reference (Python 2 + TensorFlow 1.x) cannot be imported or run in this image and ships no
golden vectors of its own, so these fixtures pin the ORACLE, not the
reference: they freeze the oracle's outputs on seeded inputs so that (a) a later edit of the oracle
that changes its arithmetic is caught by `pytest -m "not gpu"`, and (b) the CUDA path is checked on the
GPU box against numbers that were produced once, here, and committed.  Re-run after a deliberate
oracle change:   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tf_ops as T  # noqa: E402
from oracle import models as M  # noqa: E402
from oracle.kmeans import KMeans, random_init_idx  # noqa: E402


def cases():
    g = torch.Generator().manual_seed(1234)
    out = {}
    # --- adaptive front end: analysis (conv SAME + max/argmax), synthesis (unpool + transposed conv)
    x = torch.randn(3, 1024, generator=g) * 0.05
    filt = torch.randn(64, 16, generator=g) / 8.0
    X = T.conv2d_same_1d(x, filt, 1)
    y, am = T.max_pool_with_argmax_1d(X, 32, 32)
    out["analysis"] = dict(x=x, filt=filt, y=y, argmax=am)
    vals = torch.randn(2, y.shape[1], 16, generator=g)
    U = T.unpool(vals, am[:2], 1024, 16)
    out["synthesis"] = dict(vals=vals, argmax=am[:2], filt2=filt, out=T.conv2d_transpose_same_1d(U, filt, 1024, 1))
    # --- STFT twin
    xs = torch.randn(2, 2048, generator=g) * 0.05
    spec = T.stft(xs, 512, 256)
    out["stft"] = dict(x=xs, re=spec.real, im=spec.imag, istft=T.inverse_stft(spec, 512, 256))
    # --- BLSTM (BasicLSTMCell i,j,f,o; forget bias 1)
    B, Tt, I, H = 3, 9, 12, 10
    xb = torch.randn(B, Tt, I, generator=g) * 0.5
    ks = [torch.randn(I + H, 4 * H, generator=g) * 0.2 for _ in range(2)]
    bs = [torch.randn(4 * H, generator=g) * 0.1 for _ in range(2)]
    out["blstm"] = dict(x=xb, kf=ks[0], bf=bs[0], kb=ks[1], bb=bs[1], y=T.blstm(xb, ks[0], bs[0], ks[1], bs[1]))
    # --- DPCL loss and its gradient
    V = T.l2_normalize(torch.randn(2, 5, 33, 40, generator=g), 3).requires_grad_(True)
    lab = torch.randint(0, 2, (2, 5, 33), generator=g)
    Y = torch.nn.functional.one_hot(lab, 2).float()
    cost = M.dpcl_cost(V, Y)
    (dV,) = torch.autograd.grad(cost, V)
    out["dpcl"] = dict(V=V.detach(), labels=lab.to(torch.uint8), cost=cost.detach().reshape(1), dV=dV)
    # --- k-means hard assignments on margin-safe blobs
    centers = torch.randn(2, 3, 40, generator=g) * 4.0
    assign = torch.randint(0, 3, (2, 600), generator=g)
    Xk = centers[torch.arange(2)[:, None], assign] + 0.05 * torch.randn(2, 600, 40, generator=g)
    init = random_init_idx(2 * 4, 600, 3, np.random.RandomState(7))
    km = KMeans(nb_clusters=3, nb_tries=4, nb_iterations=6)
    cent, labels = km.fit(Xk, init_idx=init)
    out["kmeans"] = dict(X=Xk, init_idx=torch.as_tensor(init), labels=labels.to(torch.int32), centroids=cent)
    # --- SURVEY 8(f) rows (own generator: the fixtures above keep their values)
    g2 = torch.Generator().manual_seed(4321)
    xs3 = torch.randn(2, 3, 800, generator=g2) * 0.1
    est3 = (xs3[:, [1, 2, 0]] + 0.02 * torch.randn(2, 3, 800, generator=g2)).requires_grad_(True)
    cft = M.cost_finetuning(xs3, est3)
    (dest,) = torch.autograd.grad(cft, est3)
    out["finetune"] = dict(x_non_mix=xs3, est=est3.detach(), cost=cft.detach().reshape(1), dest=dest)
    from oracle import bss_eval as BE
    refb = torch.randn(2, 2, 1800, generator=g2, dtype=torch.float64)
    estb = torch.stack([0.8 * refb[:, 1] + 0.3 * refb[:, 0], refb[:, 0] - 0.1 * refb[:, 1]], 1) \
        + 0.05 * torch.randn(2, 2, 1800, generator=g2, dtype=torch.float64)
    res = [BE.bss_eval_sources(refb[b].numpy(), estb[b].numpy()) for b in range(2)]
    out["bss_eval"] = dict(ref=refb, est=estb, sdr=torch.tensor(np.stack([r[0] for r in res])),
                           sir=torch.tensor(np.stack([r[1] for r in res])), sar=torch.tensor(np.stack([r[2] for r in res])),
                           perm=torch.tensor(np.stack([r[3] for r in res])))
    return out


def main():
    for name, d in cases().items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v.detach().cpu().numpy()) for k, v in d.items()})
        print(name, {k: tuple(v.shape) for k, v in d.items()})


if __name__ == "__main__":
    main()
