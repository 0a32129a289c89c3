#!/usr/bin/env python
"""In-situ kernel times of the bench training step (CUPTI through torch.profiler: concurrent, warm-cache, full clocks --
unlike the serialised cold-cache ncu launch list).  Diagnostics only; prints ms/step per kernel name.
    python tools/step_profile.py [--batch 128] [--steps 3]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amss_b200  # noqa: E402,F401
import bench  # noqa: E402
from amss_b200 import models, synth, trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    t = trainer.Front_Separator_Trainer(models.DPCL, precision=a.precision, **bench.CFG)
    stream = synth.SyntheticStream(a.batch, bench.CFG["nb_speakers"], bench.L_SAMPLES, seed=42, rank=0, pool=2)
    batches = [[torch.as_tensor(x).cuda() for x in next(stream)] for _ in range(2)]
    for i in range(3):
        t.train_step(*batches[i % 2])
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(a.steps):
            t.train_step(*batches[i % 2])
        torch.cuda.synchronize()
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            tot[ev.name] += ev.device_time
            cnt[ev.name] += 1
    total = sum(tot.values())
    print(f"sum of kernel time: {total / a.steps / 1e3:.3f} ms/step")
    for name, us in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
        print(f"{us / a.steps / 1e3:8.3f} ms/step  {cnt[name] / a.steps:6.1f} launches/step  {name[:110]}")
    # the individual launches of the GEMM kernel, in order (which shape is slow?)
    g = [ev.device_time for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and "gemm_tc" in ev.name]
    per = len(g) // a.steps
    print("gemm_tc launches of the last step (us):", " ".join(f"{x:.0f}" for x in g[-per:]))


if __name__ == "__main__":
    main()
