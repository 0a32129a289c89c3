#!/usr/bin/env python
"""Golden fixture produced by THE REFERENCE ITSELF (not by the oracle): bss_eval_reference.npz.

The numpy part of the reference's metric module (utils/bss_eval.py:1-371) runs in this container once its two
unusable imports are dropped (oracle/make_ref.py writes that copy to the git-ignored oracle/_ref/).  This script feeds
it (a) the reference's own `__main__` demo inputs (utils/bss_eval.py:753-760: np.random.seed(0), two sinusoids,
reversed + 5*randn), (b) the seeded batch of tests/golden/bss_eval.npz and (c) a 3-source case, and stores inputs and
outputs (the demo inputs are regenerated from their seed by the test: demo_inputs()).  tests/test_bss_eval.py then pins oracle/bss_eval.py (CPU run) and amss_b200.bss_eval (CPU and `-m gpu`) to the
reference's numbers; /root/reference is not needed at test time.

    python oracle/make_ref.py && python tests/golden/make_bss_eval_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import make_ref  # noqa: E402


def demo_inputs():
    """utils/bss_eval.py:753-760, verbatim arithmetic."""
    np.random.seed(0)
    ts = np.linspace(0, 5, 10000)
    srcs = np.array([np.sin(ts * 600), np.cos(320 * ts + 0.01)])
    recons = srcs[::-1] + np.random.randn(*srcs.shape) * 5
    return srcs, recons


def main():
    make_ref.make()
    ref_mod = make_ref.load()
    assert ref_mod is not None, "oracle/_ref/bss_eval_ref.py missing: /root/reference not available"
    out = {}
    srcs, recons = demo_inputs()
    r = ref_mod.bss_eval_sources(srcs, recons)
    out.update(demo_sdr=r[0], demo_sir=r[1], demo_sar=r[2], demo_perm=r[3])
    r = ref_mod.bss_eval_sources(srcs, recons, compute_permutation=False)
    out.update(demo_noperm_sdr=r[0], demo_noperm_sir=r[1], demo_noperm_sar=r[2])
    with np.load(os.path.join(HERE, "bss_eval.npz")) as z:
        ref, est = z["ref"], z["est"]
    res = [ref_mod.bss_eval_sources(ref[b], est[b]) for b in range(ref.shape[0])]
    out.update(batch_ref=ref, batch_est=est, batch_sdr=np.stack([x[0] for x in res]), batch_sir=np.stack([x[1] for x in res]),
               batch_sar=np.stack([x[2] for x in res]), batch_perm=np.stack([x[3] for x in res]))
    rng = np.random.RandomState(11)
    ref3 = rng.randn(3, 1200)
    est3 = ref3[[2, 0, 1]] + 0.3 * ref3[[0, 1, 2]] + 0.1 * rng.randn(3, 1200)
    r = ref_mod.bss_eval_sources(ref3, est3)
    out.update(s3_ref=ref3, s3_est=est3, s3_sdr=r[0], s3_sir=r[1], s3_sar=r[2], s3_perm=r[3])
    np.savez_compressed(os.path.join(HERE, "bss_eval_reference.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items() if not k.endswith(("ref", "est"))})
    print("demo:", out["demo_sdr"], out["demo_sir"], out["demo_sar"], out["demo_perm"])


if __name__ == "__main__":
    main()
