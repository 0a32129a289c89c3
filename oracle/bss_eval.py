"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the reference's BSS-eval metric.

The reference vendors mir_eval's `bss_eval_sources` (utils/bss_eval.py:156-370; Vincent et al. 2006, section III.B,
512-tap time-invariant distortion filters) and calls it per mixture from its evaluation script
(experiments/evaluation/eval.py:48-73).  Restated here with numpy / scipy, loop for loop, as the checker for
amss_b200.bss_eval.  PARITY PINNED: the numpy part of the reference module runs in the build
container once its two unusable imports are dropped (oracle/make_ref.py -> oracle/_ref/bss_eval_ref.py, git-ignored);
tests/golden/bss_eval_reference.npz holds the reference's own outputs (its __main__ demo, utils/bss_eval.py:753-760, a
seeded batch and a 3-source case; generator tests/golden/make_bss_eval_golden.py) and tests/test_bss_eval.py checks this
restatement and the product against them to 1e-9 / 1e-7 dB.
"""
import itertools

import numpy as np
import scipy.fft
import scipy.linalg
import scipy.signal

FLEN = 512          # utils/bss_eval.py:236, :257 -- the filter length is hard-coded at both call sites


def project(refs, est, flen=FLEN):
    """utils/bss_eval.py:291-337: least-squares projection of `est` [L] on the span of refs [n,L] delayed by 0..flen-1."""
    n, L = refs.shape
    refs = np.hstack([refs, np.zeros((n, flen - 1))])
    est = np.hstack([est, np.zeros(flen - 1)])
    nfft = int(2 ** np.ceil(np.log2(L + flen - 1.0)))
    sf = scipy.fft.fft(refs, n=nfft, axis=1)
    sef = scipy.fft.fft(est, n=nfft)
    G = np.zeros((n * flen, n * flen))
    for i in range(n):
        for j in range(n):
            acf = np.real(scipy.fft.ifft(sf[i] * np.conj(sf[j])))
            blk = scipy.linalg.toeplitz(np.hstack([acf[0], acf[-1:-flen:-1]]), r=acf[:flen])
            G[i * flen:(i + 1) * flen, j * flen:(j + 1) * flen] = blk
            G[j * flen:(j + 1) * flen, i * flen:(i + 1) * flen] = blk.T
    D = np.zeros(n * flen)
    for i in range(n):
        xc = np.real(scipy.fft.ifft(sf[i] * np.conj(sef)))
        D[i * flen:(i + 1) * flen] = np.hstack([xc[0], xc[-1:-flen:-1]])
    try:
        C = np.linalg.solve(G, D).reshape(flen, n, order="F")
    except np.linalg.LinAlgError:
        C = np.linalg.lstsq(G, D, rcond=None)[0].reshape(flen, n, order="F")
    out = np.zeros(L + flen - 1)
    for i in range(n):
        out += scipy.signal.fftconvolve(C[:, i], refs[i])[:L + flen - 1]
    return out


def decomposition(refs, est, j, flen=FLEN):
    """utils/bss_eval.py:266-289: (s_true, e_spat, e_interf, e_artif)."""
    L = est.size
    s_true = np.hstack([refs[j], np.zeros(flen - 1)])
    e_spat = project(refs[j:j + 1], est, flen) - s_true
    e_interf = project(refs, est, flen) - s_true - e_spat
    e_artif = -s_true - e_spat - e_interf
    e_artif[:L] += est
    return s_true, e_spat, e_interf, e_artif


def _db(num, den):
    return np.inf if den == 0 else 10.0 * np.log10(num / den)     # utils/bss_eval.py:361-370


def criteria(s_true, e_spat, e_interf, e_artif):
    """utils/bss_eval.py:339-359."""
    s_filt = s_true + e_spat
    return (_db(np.sum(s_filt ** 2), np.sum((e_interf + e_artif) ** 2)), _db(np.sum(s_filt ** 2), np.sum(e_interf ** 2)),
            _db(np.sum((s_filt + e_interf) ** 2), np.sum(e_artif ** 2)))


def bss_eval_sources(refs, ests, compute_permutation=True, flen=FLEN):
    """utils/bss_eval.py:156-264: refs, ests [S,L] -> (sdr[S], sir[S], sar[S], perm[S]); the permutation maximises the
    mean SIR."""
    S = ests.shape[0]
    if not compute_permutation:
        vals = np.array([criteria(*decomposition(refs, ests[j], j, flen)) for j in range(S)])
        return vals[:, 0], vals[:, 1], vals[:, 2], np.arange(S)
    sdr, sir, sar = (np.empty((S, S)) for _ in range(3))
    for je in range(S):
        for jt in range(S):
            sdr[je, jt], sir[je, jt], sar[je, jt] = criteria(*decomposition(refs, ests[je], jt, flen))
    perms = list(itertools.permutations(range(S)))
    cols = np.arange(S)
    best = perms[int(np.argmax([np.mean(sir[list(p), cols]) for p in perms]))]
    idx = (list(best), cols)
    return sdr[idx], sir[idx], sar[idx], np.asarray(best)
