"""BSS-eval metric: the batched float64 torch implementation against the oracle's loop-for-loop
restatement of the reference's vendored mir_eval code (utils/bss_eval.py:156-370), on CPU tensors here and on the GPU in
the `-m gpu` run, plus two properties of the metric itself."""
import numpy as np
import pytest
import torch

from oracle import bss_eval as O


def _case(seed, B=2, S=2, L=2500):
    rng = np.random.RandomState(seed)
    ref = rng.randn(B, S, L)
    mixing = np.eye(S) + 0.2 * rng.rand(S, S)
    est = np.einsum("es,bsl->bel", mixing, ref)[:, ::-1].copy() + 0.05 * rng.randn(B, S, L)   # estimates in reverse order
    return ref, est


def _check(device):
    from amss_b200 import bss_eval as G
    for seed, S in ((0, 2), (1, 3)):
        ref, est = _case(seed, S=S, L=2500 if S == 2 else 1500)
        sdr, sir, sar, perm = G.bss_eval_sources(torch.tensor(ref, device=device), torch.tensor(est, device=device))
        for b in range(ref.shape[0]):
            o_sdr, o_sir, o_sar, o_perm = O.bss_eval_sources(ref[b], est[b])
            assert np.array_equal(perm[b].cpu().numpy(), o_perm)
            for got, want in ((sdr, o_sdr), (sir, o_sir), (sar, o_sar)):
                assert np.abs(got[b].cpu().numpy() - want).max() < 1e-8
        assert np.array_equal(perm[0].cpu().numpy(), np.arange(S)[::-1])          # the reversed order is recovered
        s2 = G.bss_eval_sources(torch.tensor(ref, device=device), torch.tensor(est, device=device), compute_permutation=False)
        o2 = O.bss_eval_sources(ref[0], est[0], compute_permutation=False)
        assert np.abs(s2[0][0].cpu().numpy() - o2[0]).max() < 1e-8


def test_bss_eval_matches_oracle_cpu():
    _check("cpu")


def test_bss_eval_properties():
    from amss_b200 import bss_eval as G
    ref, est = _case(3)
    r, e = torch.tensor(ref), torch.tensor(est)
    base = G.bss_eval_sources(r, e)
    scaled = G.bss_eval_sources(r, 3.0 * e)                       # the 512-tap filter absorbs any gain
    for a, b in zip(base[:3], scaled[:3]):
        assert float((a - b).abs().max()) < 1e-8
    r0 = r.clone()
    r0[..., -16:] = 0.0                                           # (so that the delayed copy loses no samples)
    delayed = torch.nn.functional.pad(r0, (7, 0))[..., :r.shape[-1]]           # a 7-sample delay is an allowed distortion
    sdr, sir, sar, _ = G.bss_eval_sources(r0, delayed)
    assert float(sdr.min()) > 60.0


@pytest.mark.gpu
def test_bss_eval_matches_oracle_gpu():
    _check("cuda")
