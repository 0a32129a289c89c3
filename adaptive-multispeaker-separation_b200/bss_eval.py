"""BSS-eval SDR / SIR / SAR as a batched device metric.

The reference scores its separations with mir_eval's `bss_eval_sources` (utils/bss_eval.py:156-370), one mixture at a
time in numpy, from experiments/evaluation/eval.py:48-73.  Mixtures are independent, so the whole evaluation batch runs
at once in float64 on the GPU: the auto- / cross-correlations through cuFFT, the Toeplitz systems
((S*512) x (S*512) for the all-sources projection, 512 x 512 for each single-source projection) assembled by one index
gather and solved by one batched factorisation, the projections by FFT convolution.  This is an evaluation metric, not
part of the training step: it is composed from torch's device FFT / solver calls (library code), there is no kernel of
ours in it; the decomposition algebra is folded so that every estimate needs only two projections:

    s_filt = P_j est,   e_interf = P_all est - P_j est,   e_artif = est - P_all est
    SDR = |s_filt|^2 / |est - s_filt|^2,  SIR = |s_filt|^2 / |e_interf|^2,  SAR = |P_all est|^2 / |e_artif|^2
"""
import itertools
import math

import torch

FLEN = 512      # utils/bss_eval.py:236, :257


def _next_pow2(n):
    return 1 << int(math.ceil(math.log2(n)))


@torch.no_grad()
def bss_eval_sources(reference_sources, estimated_sources, compute_permutation=True, flen=FLEN):
    """reference_sources, estimated_sources [B,S,L] (any float dtype, CUDA or CPU tensors) ->
    (sdr [B,S], sir [B,S], sar [B,S], perm int64 [B,S]) in float64; perm[b, j] = index of the estimate matched to true
    source j (mean-SIR criterion, as the reference), the identity if compute_permutation is False."""
    ref = reference_sources.to(torch.float64)
    est = estimated_sources.to(torch.float64)
    B, S, L = ref.shape
    dev = ref.device
    Lp = L + flen - 1
    nfft = _next_pow2(Lp)
    sf = torch.fft.fft(ref, n=nfft, dim=-1)                       # zero padding == the reference's hstack of zeros
    sef = torch.fft.fft(est, n=nfft, dim=-1)
    # acf[b,i,j,k] = sum_t ref_i[t+k] ref_j[t];  G block (i,j)[a,c] = acf[i,j,(c-a) mod nfft]
    acf = torch.fft.ifft(sf.unsqueeze(2) * sf.conj().unsqueeze(1), dim=-1).real
    a = torch.arange(flen, device=dev)
    lag = (a.view(1, flen) - a.view(flen, 1)) % nfft              # [a,c] -> (c - a) mod nfft
    Gb = acf[..., lag]                                            # [B,S,S,flen,flen]
    G_all = Gb.permute(0, 1, 3, 2, 4).reshape(B, S * flen, S * flen)
    # D[b,i,e,a] = xcorr(ref_i, est_e)[(-a) mod nfft]
    xc = torch.fft.ifft(sf.unsqueeze(2) * sef.conj().unsqueeze(1), dim=-1).real          # [B,S(i),S(e),nfft]
    D = xc[..., (-a) % nfft]                                      # [B,S,S,flen]
    D_all = D.permute(0, 1, 3, 2).reshape(B, S * flen, S)         # rhs columns = estimates
    C_all = torch.linalg.solve(G_all, D_all).reshape(B, S, flen, S)                      # [b,i,tap,e]
    G_one = torch.diagonal(Gb, dim1=1, dim2=2).permute(0, 3, 1, 2)                       # [B,S(j),flen,flen]
    C_one = torch.linalg.solve(G_one, D.permute(0, 1, 3, 2))                             # [B,S(j),flen,S(e)]
    # projections by FFT convolution of the filters with the (padded) references
    ncv = _next_pow2(Lp + flen - 1)
    rf = torch.fft.rfft(ref, n=ncv, dim=-1)                                              # [B,S,F]
    p_all = torch.fft.irfft((torch.fft.rfft(C_all.permute(0, 3, 1, 2), n=ncv, dim=-1) * rf.unsqueeze(1)).sum(2),
                            n=ncv, dim=-1)[..., :Lp]                                     # [B,e,Lp]
    p_one = torch.fft.irfft(torch.fft.rfft(C_one.permute(0, 3, 1, 2), n=ncv, dim=-1) * rf.unsqueeze(1),
                            n=ncv, dim=-1)[..., :Lp]                                     # [B,e,j,Lp]
    estp = torch.nn.functional.pad(est, (0, flen - 1))                                   # [B,e,Lp]
    s_filt = p_one
    e_interf = p_all.unsqueeze(2) - p_one
    e_artif = (estp - p_all).unsqueeze(2)

    def db(num, den):
        return torch.where(den == 0, torch.full_like(num, float("inf")), 10.0 * torch.log10(num / den))

    sdr = db((s_filt ** 2).sum(-1), ((e_interf + e_artif) ** 2).sum(-1))                 # [B,e,j]
    sir = db((s_filt ** 2).sum(-1), (e_interf ** 2).sum(-1))
    sar = db(((s_filt + e_interf) ** 2).sum(-1), (e_artif ** 2).sum(-1).expand(B, S, S))
    cols = torch.arange(S, device=dev)
    if not compute_permutation:
        perm = cols.expand(B, S).clone()
    else:
        perms = torch.tensor(list(itertools.permutations(range(S))), device=dev)         # [P,S]
        mean_sir = sir[:, perms, cols].mean(-1)                                          # [B,P]
        perm = perms[mean_sir.argmax(1)]                                                 # first maximum, like np.argmax
    bidx = torch.arange(B, device=dev).view(B, 1)
    return sdr[bidx, perm, cols], sir[bidx, perm, cols], sar[bidx, perm, cols], perm
