"""GPU parity tests: every kernel family, called through the C ABI (ctypes -> libamss_b200.so),
against the CPU oracle (oracle/) on the same seeded inputs.

Tolerances: floating point within 1e-3 relative (north_star) -- written per test as
`rel(a, b) = max|a-b| / max|b|`; integer / index outputs bit-exact on margin-safe inputs
(near-ties are excluded explicitly and the excluded fraction is bounded)."""
import math

import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from oracle import models as M
from oracle.kmeans import KMeans as OracleKMeans, random_init_idx
from oracle.amsgrad import AMSGrad as OracleAMSGrad

pytestmark = pytest.mark.gpu

REL = 1e-3


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def ops():
    import amss_b200  # noqa: F401
    from amss_b200 import ops as o
    return o


def dev(x):
    return torch.as_tensor(x).cuda().contiguous()


# ------------------------------------------------------------------------------------------ gemm
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("shape", [(37, 53, 29), (128, 128, 16), (300, 1200, 257), (1, 7, 3)])
def test_gemm_fp32(ops, ta, tb, shape):
    Mm, N, K = shape
    g = torch.Generator().manual_seed(1)
    A = torch.randn((K, Mm) if ta else (Mm, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double() + bias.double()
    out = ops.gemm(dev(A), dev(B), dev(bias), bool(ta), bool(tb))
    assert rel(out, ref) < 1e-5


def test_gemm_strided_accumulate_swap(ops):
    g = torch.Generator().manual_seed(2)
    Tt, Bb, C, N = 5, 3, 20, 33
    big = torch.randn(Tt * Bb, 2 * C, generator=g)
    A = dev(big)[:, C:]                       # row-strided view (lda = 2C)
    W = torch.randn(C, N, generator=g)
    base = torch.randn(Tt * Bb, N, generator=g)
    out = dev(base.clone())
    ops.gemm(A, dev(W), None, out=out, accumulate=True)
    ref = base.double() + big[:, C:].double() @ W.double()
    assert rel(out, ref) < 1e-5
    # time-major rows -> batch-major rows
    out2 = ops.gemm(A, dev(W), None, out_swap=(Bb, Tt))
    ref2 = (big[:, C:].double() @ W.double()).reshape(Tt, Bb, N).transpose(0, 1).reshape(Bb * Tt, N)
    assert rel(out2, ref2) < 1e-5


# ------------------------------------------------------------------------------------------ STFT
def test_stft_matches_oracle(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 4096, generator=g) * 0.05
    ref = T.stft(x, 512, 256)
    spec, mag = ops.stft(dev(x), 512, 256)
    assert spec.shape == ref.shape == (3, 15, 257)
    assert rel(torch.view_as_real(spec), torch.view_as_real(ref)) < 1e-4
    assert rel(mag, ref.abs()) < 1e-4


@pytest.mark.parametrize("frame,hop", [(64, 16), (256, 128), (1024, 256)])
def test_stft_other_frames(ops, frame, hop):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 5000, generator=g)
    ref = T.stft(x, frame, hop)
    spec, mag = ops.stft(dev(x), frame, hop)
    assert rel(torch.view_as_real(spec), torch.view_as_real(ref)) < 1e-4


@pytest.mark.parametrize("R,Lw,frame,hop", [(40, 30000, 512, 256), (14, 5000, 64, 16), (9, 120000, 1024, 256), (36, 31232, 512, 256)])
def test_stft_run_kernels_match_oracle(ops, R, Lw, frame, hop):
    """Batches of >= 4000 frames take the multi-frame kernels (two real frames per complex FFT, window / twiddles once per
    CTA): odd and even frame counts, runs that end inside a pair, other frame sizes; spectra, magnitudes and labels."""
    g = torch.Generator().manual_seed(6)
    x = torch.randn(R, Lw, generator=g) * 0.1
    Tn = 1 + (Lw - frame) // hop
    assert Tn * R >= 4000
    ref = T.stft(x, frame, hop)
    spec, mag = ops.stft(dev(x), frame, hop)
    assert spec.shape == ref.shape
    assert rel(torch.view_as_real(spec), torch.view_as_real(ref)) < 1e-4
    assert rel(mag, ref.abs()) < 1e-4
    S = 3
    B = R // S
    nm = x[:B * S].reshape(B, S, Lw)
    if Tn * B >= 4000:
        refm = ref[:B * S].abs().reshape(B, S, Tn, -1).permute(0, 2, 3, 1)       # [B,T,F,S]
        labels, magn = ops.stft_labels(dev(nm), frame, hop, want_mag=True)
        assert rel(magn, refm) < 1e-4
        srt = refm.sort(-1).values
        margin = (srt[..., -1] - srt[..., -2]) > 1e-4 * srt[..., -1].clamp_min(1e-6)
        assert torch.equal(labels.cpu().long()[margin], refm.argmax(-1)[margin])


def test_stft_labels_bit_exact_off_ties(ops):
    mix, nm, _ = M.synthetic_mixtures(2, 3, 8192, seed=5)
    pre = M.separator_preprocessing(torch.tensor(mix), torch.tensor(nm), 512, 256, 1.0, 0.0)
    labels, mag = ops.stft_labels(dev(nm), 512, 256, want_mag=True)
    assert rel(mag, pre["X_non_mix"]) < 1e-4
    srt = pre["X_non_mix"].sort(-1).values
    margin = (srt[..., -1] - srt[..., -2]) > 1e-4 * srt[..., -1].clamp_min(1e-6)
    assert margin.float().mean() > 0.9
    assert torch.equal(labels.cpu().long()[margin], pre["argmax"][margin])


def test_istft_masked_matches_oracle(ops):
    mix, nm, _ = M.synthetic_mixtures(2, 2, 8192, seed=6)
    pre = M.separator_preprocessing(torch.tensor(mix), torch.tensor(nm), 512, 256, 1.0, 0.0)
    B, Tt, Fb = pre["X"].shape
    labels = pre["argmax"].reshape(B, Tt * Fb)
    sep, masks = M.separate(torch.zeros(B, Tt, Fb, 1), pre["X"], lambda V: labels, 2)
    ref = M.postprocessing(sep, pre["stfts"], 2, 512, 256)
    spec, _ = ops.stft(dev(mix), 512, 256)
    out = ops.istft_masked(spec, 2, 512, 256, labels=dev(labels.to(torch.int32)))
    assert out.shape == ref.shape
    assert rel(out, ref) < 1e-4
    # soft masks
    g = torch.Generator().manual_seed(7)
    soft = torch.softmax(torch.randn(B, Tt * Fb, 2, generator=g), -1)
    sep2 = (pre["X"].reshape(B, -1, 1) * soft).reshape(B, Tt, Fb, 2).permute(0, 3, 1, 2).reshape(B * 2, Tt, Fb)
    ref2 = M.postprocessing(sep2, pre["stfts"], 2, 512, 256)
    out2 = ops.istft_masked(spec, 2, 512, 256, masks=dev(soft))
    assert rel(out2, ref2) < 1e-4


@pytest.mark.parametrize("Lw,frame,hop,S", [(8192, 512, 256, 2), (5000, 256, 64, 3), (64000, 512, 256, 1), (700, 512, 128, 2)])
def test_istft_run_kernel_matches_per_block_kernel(ops, Lw, frame, hop, S):
    """The inverse STFT that evaluates every frame once (runs of 16 output blocks per CTA, ring of N / hop block accumulators)
    against the one-CTA-per-block kernel (AMSS_ISTFT_PER_BLOCK=1): hop = frame / 2, / 4, / 8, block counts that are not a
    multiple of the run length, a signal with fewer frames than one run; hard labels and soft masks."""
    import os
    g = torch.Generator().manual_seed(11 + Lw)
    x = dev(torch.randn(2, Lw, generator=g) * 0.1)
    spec, _ = ops.stft(x, frame, hop)
    B, Tt, Fb = spec.shape[0], spec.shape[1], spec.shape[2]
    soft = dev(torch.softmax(torch.randn(B, Tt * Fb, S, generator=g), -1))
    lab = dev(torch.randint(0, S, (B, Tt * Fb), generator=g).to(torch.int32))
    fast = [ops.istft_masked(spec, S, frame, hop, masks=soft), ops.istft_masked(spec, S, frame, hop, labels=lab)]
    os.environ["AMSS_ISTFT_PER_BLOCK"] = "1"
    try:
        slow = [ops.istft_masked(spec, S, frame, hop, masks=soft), ops.istft_masked(spec, S, frame, hop, labels=lab)]
    finally:
        del os.environ["AMSS_ISTFT_PER_BLOCK"]
    for a, b in zip(fast, slow):
        assert a.shape == b.shape and rel(a, b) < 1e-5


def test_stft_istft_round_trip_full_size(ops):
    """Size-independent property at the BASELINE size (L=64000): all-ones masks give back the
    mixture on [hop, L-hop) (the TF inverse window does not reconstruct the edges)."""
    mix, nm, _ = M.synthetic_mixtures(2, 2, 64000, seed=8)
    spec, _ = ops.stft(dev(mix), 512, 256)
    ones = torch.ones(2, spec.shape[1] * spec.shape[2], 1, device="cuda")
    out = ops.istft_masked(spec, 1, 512, 256, masks=ones)
    ref = torch.tensor(mix)
    assert rel(out[:, 0, 256:-256], ref[:, 256:-256]) < 1e-4


# ------------------------------------------------------------------------------------------ BLSTM
def _blstm_case(ops, B, Tt, I, H, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Tt, I, generator=g, dtype=torch.float64) * 0.5
    ks = [(torch.rand(I + H, 4 * H, generator=g, dtype=torch.float64) * 2 - 1) * math.sqrt(6.0 / (I + 5 * H))
          for _ in range(2)]
    bs = [torch.randn(4 * H, generator=g, dtype=torch.float64) * 0.1 for _ in range(2)]
    dy = torch.randn(B, Tt, 2 * H, generator=g, dtype=torch.float64)
    leaves = [x] + ks + bs
    for t in leaves:
        t.requires_grad_(True)
    y = T.blstm(x, ks[0], bs[0], ks[1], bs[1])
    grads = torch.autograd.grad((y * dy).sum(), leaves)
    f = lambda t: dev(t.detach().float())
    x_tm = ops.transpose_01(f(x))
    y_tm, saved = ops.blstm_fwd(x_tm, f(ks[0]), f(bs[0]), f(ks[1]), f(bs[1]))
    y_gpu = ops.transpose_01(y_tm)
    assert rel(y_gpu, y) < REL
    dy_tm = ops.transpose_01(f(dy))
    dx, dkf, dbf, dkb, dbb = ops.blstm_bwd(x_tm, f(ks[0]), f(ks[1]), y_tm, dy_tm, saved)
    assert rel(ops.transpose_01(dx), grads[0]) < REL
    assert rel(dkf, grads[1]) < REL and rel(dkb, grads[2]) < REL
    assert rel(dbf, grads[3]) < REL and rel(dbb, grads[4]) < REL


@pytest.mark.parametrize("B,Tt,I,H", [(3, 7, 11, 6), (1, 1, 5, 4), (4, 20, 129, 150), (70, 5, 33, 20)])
def test_blstm_fwd_bwd_matches_oracle(ops, B, Tt, I, H):
    _blstm_case(ops, B, Tt, I, H, seed=10 + B)


def test_blstm_matches_oracle_reference_width(ops):
    # the BASELINE width (layer_size 600 -> H=300 per direction), short sequence
    _blstm_case(ops, 2, 12, 64, 300, seed=20)


# ------------------------------------------------------------------------------------------ head / losses
def test_l2norm_fwd_bwd(ops):
    g = torch.Generator().manual_seed(30)
    z = torch.randn(500, 40, generator=g, dtype=torch.float64)
    z[7] = 0.0                                    # the clamped branch (sum z^2 < 1e-12)
    z.requires_grad_(True)
    dv = torch.randn(500, 40, generator=g, dtype=torch.float64)
    v = T.l2_normalize(z, -1)
    (gz,) = torch.autograd.grad((v * dv).sum(), z)
    vg, inv = ops.l2norm_fwd(dev(z.detach().float()), 40)
    assert rel(vg, v) < 1e-5
    dz = ops.l2norm_bwd(vg, inv, dev(dv.float()), 40)
    assert rel(dz, gz) < 1e-4


def test_colsum(ops):
    g = torch.Generator().manual_seed(31)
    Z = torch.randn(1000, 77, generator=g)
    assert rel(ops.colsum(dev(Z)), Z.double().sum(0)) < 1e-5


@pytest.mark.parametrize("S", [2, 3])
def test_dpcl_loss_fwd_bwd(ops, S):
    g = torch.Generator().manual_seed(32)
    B, Tt, Fb, E = 3, 9, 33, 40
    V = T.l2_normalize(torch.randn(B, Tt, Fb, E, generator=g, dtype=torch.float64), 3).requires_grad_(True)
    lab = torch.randint(0, S, (B, Tt, Fb), generator=g)
    y = torch.nn.functional.one_hot(lab, S).double()
    cost = M.dpcl_cost(V, y)
    (gV,) = torch.autograd.grad(cost, V)
    Vg = dev(V.detach().float().reshape(B, Tt * Fb, E))
    labg = dev(lab.to(torch.uint8).reshape(B, Tt * Fb))
    loss, ws = ops.dpcl_loss_fwd(Vg, labg, S)
    assert abs(float(loss) - float(cost)) < REL * abs(float(cost))
    dV = ops.dpcl_loss_bwd(Vg, labg, S, torch.ones(1, device="cuda"), ws)
    assert rel(dV, gV.reshape(B, Tt * Fb, E)) < REL


@pytest.mark.parametrize("S,E", [(2, 40), (3, 12)])
def test_dpcl_loss_weighted_fwd_bwd(ops, S, E):
    """DPCL.cost on the weighted label matrix of --function_mask (models/network.py:381-389 -> models/dpcl.py:41-86):
    Y = w * one_hot; cost, dV, and dz through the l2_normalize Jacobian against the oracle's general-Y formula."""
    g = torch.Generator().manual_seed(34)
    B, Tt, Fb = 3, 9, 33
    z = torch.randn(B, Tt, Fb, E, generator=g, dtype=torch.float64).requires_grad_(True)
    V = T.l2_normalize(z, 3)
    lab = torch.randint(0, S, (B, Tt, Fb), generator=g)
    w = torch.rand(B, Tt, Fb, generator=g, dtype=torch.float64) * 0.9 + 0.1
    y = torch.nn.functional.one_hot(lab, S).double() * w.unsqueeze(-1)
    cost = M.dpcl_cost(V, y)
    gV, gz = torch.autograd.grad(cost, [V, z])
    vg, inv = ops.l2norm_fwd(dev(z.detach().float().reshape(-1, E)), E)
    Vg = vg.view(B, Tt * Fb, E)
    labg = dev(lab.to(torch.uint8).reshape(B, Tt * Fb))
    wg = dev(w.float().reshape(B, Tt * Fb))
    loss, ws = ops.dpcl_loss_weighted_fwd(Vg, labg, wg, S)
    assert abs(float(loss) - float(cost)) < REL * abs(float(cost))
    one = torch.ones(1, device="cuda")
    assert rel(ops.dpcl_loss_weighted_bwd(Vg, labg, wg, S, one, ws), gV.reshape(B, Tt * Fb, E)) < REL
    assert rel(ops.dpcl_loss_weighted_bwd(Vg, labg, wg, S, one, ws, inv), gz.reshape(B, Tt * Fb, E)) < REL
    # unit weights reproduce the one-hot kernels
    l1, ws1 = ops.dpcl_loss_weighted_fwd(Vg, labg, torch.ones_like(wg), S)
    l0, ws0 = ops.dpcl_loss_fwd(Vg, labg, S)
    assert abs(float(l1) - float(l0)) < 1e-6 * abs(float(l0))
    assert rel(ops.dpcl_loss_weighted_bwd(Vg, labg, torch.ones_like(wg), S, one, ws1), ops.dpcl_loss_bwd(Vg, labg, S, one, ws0)) < 1e-6


def test_l41_loss_fwd_bwd(ops):
    g = torch.Generator().manual_seed(33)
    B, Tt, Fb, E, S = 2, 7, 21, 40, 2
    emb = T.l2_normalize(torch.randn(B, Tt, Fb, E, generator=g, dtype=torch.float64), 3).requires_grad_(True)
    lab = torch.randint(0, S, (B, Tt, Fb), generator=g)
    y = torch.nn.functional.one_hot(lab, S).double() * 2 - 1
    spk = T.l2_normalize(torch.randn(B, S, E, generator=g, dtype=torch.float64), -1).requires_grad_(True)
    dot = (spk[:, None, None, :, :] * emb[:, :, :, None, :]).sum(4)
    cost = (-torch.log(torch.sigmoid(y * dot))).mean(3).mean(0).mean()
    gE, gS = torch.autograd.grad(cost, [emb, spk])
    embg = dev(emb.detach().float().reshape(B, Tt * Fb, E))
    labg = dev(lab.to(torch.uint8).reshape(B, Tt * Fb))
    spkg = dev(spk.detach().float())
    loss = ops.l41_loss_fwd(embg, labg, spkg)
    assert abs(float(loss) - float(cost)) < REL * abs(float(cost))
    demb, dspk = ops.l41_loss_bwd(embg, labg, spkg, torch.ones(1, device="cuda"))
    assert rel(demb, gE.reshape(B, Tt * Fb, E)) < REL
    assert rel(dspk, gS) < REL


# ------------------------------------------------------------------------------------------ filterbank
@pytest.mark.parametrize("L,W,N,pool,hop", [(2048, 64, 32, 64, 64), (3000, 100, 70, 128, 64), (4096, 1024, 256, 256, 256),
                                            (1500, 33, 8, 300, 100)])
def test_filterbank_analysis_max(ops, L, W, N, pool, hop):
    g = torch.Generator().manual_seed(40)
    x = torch.randn(3, L, generator=g) * 0.1
    filt = torch.randn(W, N, generator=g) / math.sqrt(W)
    X = T.conv2d_same_1d(x.double(), filt.double(), 1)
    y_ref, am_ref = T.max_pool_with_argmax_1d(X, pool, hop)
    y, am = ops.filterbank_analysis(dev(x), dev(filt), pool, hop, ops.AMSS_POOL_MAX)
    assert y.shape == y_ref.shape
    assert rel(y, y_ref) < 1e-4
    am = am.cpu()
    t = (am // N)
    n = (am % N)
    assert torch.equal(n, torch.arange(N).view(1, 1, N).expand_as(n))
    # every picked position lies in its window and is a near-maximiser of the oracle's X
    tp = torch.arange(y.shape[1]).view(1, -1, 1)
    assert bool(((t >= tp * hop) & (t < tp * hop + pool)).all())
    picked = torch.gather(X, 1, t)
    assert float((y_ref - picked).abs().max()) < 1e-5 * float(y_ref.abs().max())
    assert float((am == am_ref).float().mean()) > 0.999


def test_filterbank_analysis_avg_and_stride(ops):
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, 2100, generator=g) * 0.1
    filt = torch.randn(50, 24, generator=g) / 7.0
    X = T.conv2d_same_1d(x.double(), filt.double(), 1)
    y, _ = ops.filterbank_analysis(dev(x), dev(filt), 64, 64, ops.AMSS_POOL_AVG)
    assert rel(y, T.avg_pool_1d(X, 64)) < 1e-4
    ys, _ = ops.filterbank_analysis(dev(x), dev(filt), 64, 48, ops.AMSS_POOL_STRIDE)
    assert rel(ys, T.conv2d_same_1d(x.double(), filt.double(), 48)) < 1e-4


def test_make_filter_fwd_bwd(ops):
    g = torch.Generator().manual_seed(42)
    w = torch.randn(64, generator=g, dtype=torch.float64).requires_grad_(True)
    b = torch.randn(64, 16, generator=g, dtype=torch.float64).requires_grad_(True)
    df = torch.randn(64, 16, generator=g, dtype=torch.float64)
    f = w.abs().unsqueeze(1) * b
    gw, gb = torch.autograd.grad((f * df).sum(), [w, b])
    fg = ops.make_filter(dev(w.detach().float()), dev(b.detach().float()))
    assert rel(fg, f) < 1e-6
    dw, db = ops.make_filter_bwd(dev(w.detach().float()), dev(b.detach().float()), dev(df.float()))
    assert rel(dw, gw) < 1e-5 and rel(db, gb) < 1e-5


@pytest.mark.parametrize("L,W,N,pool,hop,B,S", [(2048, 64, 32, 64, 64, 2, 2), (3000, 100, 24, 128, 64, 1, 3),
                                                (4096, 1024, 256, 256, 256, 1, 2)])
def test_filterbank_synthesis_fwd_bwd_and_analysis_bwd(ops, L, W, N, pool, hop, B, S):
    g = torch.Generator().manual_seed(43)
    x = torch.randn(B * (S + 1), L, generator=g, dtype=torch.float64) * 0.1
    filt = (torch.randn(W, N, generator=g, dtype=torch.float64) / math.sqrt(W)).requires_grad_(True)
    filt2 = (torch.randn(W, N, generator=g, dtype=torch.float64) / math.sqrt(W)).requires_grad_(True)
    X = T.conv2d_same_1d(x, filt, 1)
    y, am = T.max_pool_with_argmax_1d(X, pool, hop)
    Tp = y.shape[1]
    vals = y[B:].detach().clone().requires_grad_(True)          # 'mask' separation == non-mix rows
    am_mix = am[:B].unsqueeze(1).repeat(1, S, 1, 1).reshape(B * S, Tp, N)
    U = T.unpool(vals, am_mix, L, N)
    out = T.conv2d_transpose_same_1d(U, filt2, L, 1)
    dout = torch.randn(B * S, L, generator=g, dtype=torch.float64)
    gvals, gf2 = torch.autograd.grad((out * dout).sum(), [vals, filt2])
    f = lambda t: dev(t.detach().float())
    # use the ORACLE's argmax so that the comparison is not perturbed by near-ties
    am_g = dev(am[:B])
    og = ops.filterbank_synthesis(f(vals), am_g, f(filt2), B, S, L, pool, hop)
    assert rel(og, out) < 1e-4
    dv, df2 = ops.filterbank_synthesis_bwd(f(dout), f(vals), am_g, f(filt2), B, S)
    assert rel(dv, gvals) < 1e-4
    assert rel(df2, gf2) < 1e-4
    # analysis backward w.r.t. the filter, through the arg-max
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (gf,) = torch.autograd.grad((y * dy).sum(), filt)
    dfilt = ops.filterbank_analysis_bwd(f(x), f(dy), dev(am), W)
    assert rel(dfilt, gf) < 1e-4


@pytest.mark.parametrize("L,N,B,S", [(5000, 40, 2, 2), (2048, 16, 3, 1), (13000, 256, 1, 3), (700, 24, 2, 4)])
def test_synthesis_filter_stationary_kernel_matches_oracle_and_gather(ops, L, N, B, S):
    """The filter-stationary overlap-add (W = 1024 taps, pool = hop = 256: the reference geometry) against the oracle's
    unpool + conv2d_transpose (utils/ops.py:94-120, adapt.py:241-243) and against the gather kernel
    (AMSS_SYNTHESIS_GATHER=1): ragged last block (L % 256 != 0), a filter count that does not fill the last group of 16,
    one to four rows per mixture, signals shorter than the filter."""
    import os
    W, pool = 1024, 256
    g = torch.Generator().manual_seed(47 + L)
    x = torch.randn(B * (S + 1), L, generator=g, dtype=torch.float64) * 0.1
    filt = torch.randn(W, N, generator=g, dtype=torch.float64) / math.sqrt(W)
    filt2 = torch.randn(W, N, generator=g, dtype=torch.float64) / math.sqrt(W)
    y, am = T.max_pool_with_argmax_1d(T.conv2d_same_1d(x, filt, 1), pool, pool)
    Tp = y.shape[1]
    vals = y[B:].clone()
    am_mix = am[:B].unsqueeze(1).repeat(1, S, 1, 1).reshape(B * S, Tp, N)
    out = T.conv2d_transpose_same_1d(T.unpool(vals, am_mix, L, N), filt2, L, 1)
    f = lambda t: dev(t.float())
    fast = ops.filterbank_synthesis(f(vals), dev(am[:B]), f(filt2), B, S, L, pool, pool)
    os.environ["AMSS_SYNTHESIS_GATHER"] = "1"
    try:
        slow = ops.filterbank_synthesis(f(vals), dev(am[:B]), f(filt2), B, S, L, pool, pool)
    finally:
        del os.environ["AMSS_SYNTHESIS_GATHER"]
    assert rel(fast, out) < 1e-4 and rel(slow, out) < 1e-4
    assert rel(fast, slow) < 1e-5


def test_box_sum_rows_columns_and_adjoint(ops):
    """amss_box_sum: sliding sums along rows (signals) and columns (filter banks), both directions, zero outside; the
    column adjoint used for the filter gradient of the average-pool back end."""
    g = torch.Generator().manual_seed(51)
    x = torch.randn(3, 700, generator=g, dtype=torch.float64)
    P = 32
    xp = torch.nn.functional.pad(x, (0, P - 1))
    ref = xp.unfold(1, P, 1).sum(-1) / P                                   # (1/P) sum_j x[u + j]
    assert rel(ops.box_sum(dev(x.float()), P, dir=1, scale=1.0 / P), ref) < 1e-5
    w = torch.randn(64, 24, generator=g, dtype=torch.float64)
    wp = torch.nn.functional.pad(w.t(), (P - 1, P - 1))                    # [N, W + 2(P-1)]
    wbox = wp.unfold(1, P, 1).sum(-1).t()                                  # wbox[k'] = sum_j w[k' - j], k' < W + P - 1
    got = ops.box_sum(dev(w.float()), P, dir=-1, len_out=64 + P - 1, axis=0)
    assert got.shape == (64 + P - 1, 24) and rel(got, wbox) < 1e-5
    d = torch.randn(64 + P - 1, 24, generator=g, dtype=torch.float64)
    adj = ops.box_sum(dev(d.float()), P, dir=1, len_out=64, axis=0)
    assert abs(float((wbox * d).sum()) - float((w * adj.double().cpu()).sum())) < 1e-3 * float((wbox * d).abs().sum())


def test_synthesis_is_adjoint_of_analysis_full_size(ops):
    """<A x, y> = <x, A^T y> at the BASELINE size (L=64000, W=1024, N=256, pool=hop=256), using the
    sparse structure: scatter y at the arg-max positions == synthesis with the same filter."""
    g = torch.Generator().manual_seed(44)
    L, W, N = 64000, 1024, 256
    x = dev(torch.randn(1, L, generator=g) * 0.1)
    filt = dev(torch.randn(W, N, generator=g) / 32.0)
    y, am = ops.filterbank_analysis(x, filt, 256, 256, ops.AMSS_POOL_MAX)
    c = dev(torch.randn(1, y.shape[1], N, generator=g))
    lhs = float((y.double() * c.double()).sum())          # <P A x, c> with P = arg-max selection
    back = ops.filterbank_synthesis(c, am, filt, 1, 1, L, 256, 256)
    rhs = float((x.double() * back.double()).sum())       # <x, A^T P^T c>
    assert abs(lhs - rhs) < 1e-3 * max(abs(lhs), 1.0)


def test_plugged_labels_and_wave_stats(ops):
    g = torch.Generator().manual_seed(45)
    B, S, Tp, N = 3, 2, 10, 16
    y = torch.randn(B * (S + 1), Tp, N, generator=g)
    ref = M.separator_plugged_inputs(y, B, S, 1.0, 0.0)["argmax"]
    lab = ops.plugged_labels(dev(y), B, S)
    assert torch.equal(lab.cpu().long(), ref)
    a = torch.randn(5, 3001, generator=g)
    b = torch.randn(5, 3001, generator=g)
    st = ops.wave_stats(dev(a), dev(b)).cpu().double()
    ref4 = torch.stack([(a.double() ** 2).sum(1), (b.double() ** 2).sum(1), (a.double() * b.double()).sum(1),
                        ((a.double() - b.double()) ** 2).sum(1)], 1)
    assert rel(st, ref4) < 1e-5


# ------------------------------------------------------------------------------------------ k-means
def _blobs(B, L, E, K, seed, spread=0.05):
    rng = np.random.RandomState(seed)
    X = np.zeros((B, L, E), np.float32)
    truth = np.zeros((B, L), np.int64)
    for b in range(B):
        centers = rng.randn(K, E)
        centers /= np.linalg.norm(centers, axis=1, keepdims=True)
        truth[b] = rng.randint(0, K, L)
        X[b] = centers[truth[b]] + spread * rng.randn(L, E)
    return X, truth


def _same_partition(a, b, K):
    pairs = set(zip(a.tolist(), b.tolist()))
    return len(pairs) == len({p[0] for p in pairs}) == len({p[1] for p in pairs})


@pytest.mark.parametrize("K,tries,with_silence,assign_at_end", [(2, 1, False, True), (3, 4, False, True),
                                                                (2, 3, True, True), (3, 2, True, False)])
def test_kmeans_hard_bit_exact(ops, K, tries, with_silence, assign_at_end):
    B, L, E, iters = 3, 3000, 40, 6
    X, _ = _blobs(B, L, E, K, seed=50 + K)
    rng = np.random.RandomState(60)
    idx = random_init_idx(B * tries, L, K, rng)
    latent = None
    ns = None
    if with_silence:
        latent = np.abs(rng.randn(B, L)).astype(np.float32) + 1e-3
        latent[:, ::7] *= 1e-4                                 # silent bins
    okm = OracleKMeans(K, tries, iters, True, None, 2.0, assign_at_end)
    c_ref, l_ref = okm.fit(torch.tensor(X), idx, None if latent is None else torch.tensor(latent))
    if with_silence:
        ns = ops.kmeans_silence_mask(dev(latent), 2.0)
        ref_ns = (T.log10(torch.tensor(latent).max(-1, keepdim=True).values / torch.tensor(latent)) < 2.0)
        assert torch.equal(ns.cpu().bool(), ref_ns)
    cent, labels, inertia, best = ops.kmeans_fit(dev(X), dev(idx), K, tries, iters, None, ns, True, assign_at_end)
    ok = ~torch.isnan(okm.last_inertia)
    assert torch.equal(torch.isnan(inertia.cpu()), ~ok)
    assert rel(inertia.cpu()[ok], okm.last_inertia[ok]) < 1e-4
    best = best.cpu().long()
    for b in range(B):
        if int(best[b]) == int(okm.last_best[b]):
            assert rel(cent[b], c_ref[b]) < 1e-4
            assert torch.equal(labels[b].cpu(), l_ref[b])      # bit-exact hard assignments
        else:
            # two tries converged to the same optimum: inertias tie to rounding, the chosen try (and
            # so the numbering of the clusters) may differ -- the partition must still be identical
            i_ref = okm.last_inertia[b]
            assert abs(float(i_ref[best[b]] - i_ref[okm.last_best[b]])) < 1e-5 * float(i_ref[okm.last_best[b]])
            assert _same_partition(labels[b].cpu(), l_ref[b], K)


@pytest.mark.parametrize("B,L,K,tries", [(1, 100, 2, 1), (1, 128, 3, 10), (2, 129, 4, 8), (3, 257, 3, 10), (5, 1000, 2, 16),
                                         (20, 700, 3, 4), (150, 300, 3, 2), (2, 20000, 3, 10)])
def test_kmeans_tensor_core_pass_matches_simt_kernels(ops, B, L, K, tries, monkeypatch):
    """The tcgen05 k-means pass (kmeans_tc.cu: TMA tile ring, two loader groups on alternate tiles, chained passes, one
    wave of CTAs) against the fp32 SIMT kernels (AMSS_KMEANS_SIMT=1) over geometries that exercise its edges: a single
    partial tile, exact / off-by-one tile counts, fewer tiles than ring slots, more mixtures than SMs (chunks = 1; with two
    tries some mixtures end with an EMPTY cluster, whose NaN centroid must never be anybody's nearest), tries * K = 32, many
    tiles per CTA.  Same initial rows: same best try, same labels off near-ties, centroids to 1e-5."""
    E, iters = 40, 5
    X, _ = _blobs(B, L, E, K, seed=300 + B + L)
    idx = random_init_idx(B * tries, L, K, np.random.RandomState(301))
    Xd, idxd = dev(X), dev(idx)
    monkeypatch.setenv("AMSS_KMEANS_SIMT", "1")
    c0, l0, i0, b0 = ops.kmeans_fit(Xd, idxd, K, tries, iters, None, None, True, True)
    monkeypatch.setenv("AMSS_KMEANS_SIMT", "0")
    c1, l1, i1, b1 = ops.kmeans_fit(Xd, idxd, K, tries, iters, None, None, True, True)
    torch.cuda.synchronize()
    ok = ~torch.isnan(i0)
    assert torch.equal(torch.isnan(i1), ~ok)
    # (the tensor-core pass evaluates |x|^2 - 2 x.c + |c|^2, the SIMT kernels sum (x - c)^2: the inertia of tight blobs is
    # a sum of cancellations in the former)
    assert rel(i1[ok], i0[ok]) < 2e-4
    Xn = torch.tensor(X) / torch.tensor(X).norm(dim=-1, keepdim=True)
    for b in range(B):
        if int(b0[b]) != int(b1[b]):                           # two tries tie to rounding: same partition required
            assert abs(float(i0[b, b1[b]] - i0[b, b0[b]])) < 2e-4 * abs(float(i0[b, b0[b]]))
            assert _same_partition(l1[b].cpu(), l0[b].cpu(), K)
            continue
        nan0 = torch.isnan(c0[b])                              # an empty cluster's centroid is 0/0 in both paths (Kmeans_2.py:158-165)
        assert torch.equal(torch.isnan(c1[b]), nan0)
        assert rel(c1[b][~nan0], c0[b][~nan0]) < 1e-5
        d = ((Xn[b].unsqueeze(1) - c0[b].cpu().unsqueeze(0)) ** 2).sum(-1).sort(-1).values
        safe = (d[:, 1] - d[:, 0]) > 1e-4
        assert torch.equal(l1[b].cpu()[safe], l0[b].cpu()[safe])


def test_kmeans_soft_matches_oracle(ops):
    B, L, E, K, tries, iters = 2, 2000, 40, 2, 3, 5
    X, _ = _blobs(B, L, E, K, seed=70, spread=0.2)
    idx = random_init_idx(B * tries, L, K, np.random.RandomState(71))
    okm = OracleKMeans(K, tries, iters, True, 5.0, 2.0, True)
    c_ref, l_ref = okm.fit(torch.tensor(X), idx)
    cent, soft, inertia, best = ops.kmeans_fit(dev(X), dev(idx), K, tries, iters, 5.0, None, True, True)
    assert torch.equal(best.cpu().long(), okm.last_best)
    assert rel(cent, c_ref) < 1e-4
    assert rel(soft, l_ref) < 1e-4


def test_kmeans_recovers_blobs_reference_smoke(ops):
    """Mirrors the reference's own smoke block (models/Kmeans_2.py:197-219): well-separated blobs,
    4 clusters, 40 features, 10 tries x 10 iterations -> the generating partition (up to a
    permutation of the labels)."""
    B, L, E, K = 3, 1000, 40, 4
    X, truth = _blobs(B, L, E, K, seed=80, spread=0.02)
    idx = random_init_idx(B * 10, L, K, np.random.RandomState(81))
    _, labels, _, _ = ops.kmeans_fit(dev(X), dev(idx), K, 10, 10, None, None, True, True)
    labels = labels.cpu().numpy()
    for b in range(B):
        mapping = {}
        for l, t in zip(labels[b], truth[b]):
            assert mapping.setdefault(int(l), int(t)) == int(t)
        assert len(set(mapping.values())) == K


def test_kmeans_full_size_idempotent(ops):
    """BASELINE size (TF = 63993 bins, E = 40, K = 3, 10 tries): a converged fit is a fixed point
    -- re-running one more iteration from the returned centroids reproduces labels exactly."""
    B, L, E, K = 2, 63993, 40, 3
    X, truth = _blobs(B, L, E, K, seed=90, spread=0.05)
    idx = random_init_idx(B * 10, L, K, np.random.RandomState(91))
    cent, labels, inertia, best = ops.kmeans_fit(dev(X), dev(idx), K, 10, 10, None, None, True, True)
    assert bool(torch.isfinite(inertia).all())
    Xn = torch.tensor(X) / torch.tensor(X).norm(dim=-1, keepdim=True)
    d = ((Xn.unsqueeze(2) - cent.cpu().unsqueeze(1)) ** 2).sum(-1)
    srt = d.sort(-1).values
    safe = (srt[..., 1] - srt[..., 0]) > 1e-4
    assert torch.equal(d.argmin(-1)[safe].int(), labels.cpu()[safe])
    masks = ops.apply_masks(dev(np.ones((B, L), np.float32)), K, labels=labels)
    assert float(masks.reshape(B, K, L).sum(1).min()) == 1.0 and float(masks.sum()) == B * L


# ------------------------------------------------------------------------------------------ optimizer
def test_amsgrad_matches_oracle(ops):
    g = torch.Generator().manual_seed(100)
    n = 10007
    p0 = torch.randn(n, generator=g)
    params = {"p": p0.clone()}
    opt = OracleAMSGrad(params, lr=1e-3, clip=0.5)
    p = dev(p0.clone())
    m, v, vh = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g)
        opt.step({"p": grad})
        gd = dev(grad)
        fac = ops.global_norm_clip_factor(gd, 0.5)
        ops.amsgrad_step(p, gd, m, v, vh, ops.amsgrad_lr_t(1e-3, 0.9, 0.99, step), 0.9, 0.99, 1e-3, 1.0, fac)
    assert rel(p, params["p"]) < 1e-5


@pytest.mark.parametrize("kind", ["SGD", "RMSProp"])
def test_momentum_rmsprop_match_oracle(ops, kind):
    """--optimizer SGD / RMSProp (models/network.py:183-186) with the staircase decay and a global-norm clip."""
    from oracle.amsgrad import make_optimizer
    g = torch.Generator().manual_seed(101)
    n = 10007
    p0 = torch.randn(n, generator=g)
    params = {"p": p0.clone()}
    opt = make_optimizer(kind, params, 1e-2, decay_epoch=1, clip=0.5)
    p = dev(p0.clone())
    s1 = torch.ones_like(p) if kind == "RMSProp" else torch.zeros_like(p)
    s2 = torch.zeros_like(p)
    for step in range(4):
        if step == 2:
            opt.increment_epoch()                     # decay_epoch 1: the rate halves from here on
        lr = 1e-2 * (0.5 if step >= 2 else 1.0)
        grad = torch.randn(n, generator=g)
        opt.step({"p": grad})
        gd = dev(grad)
        fac = ops.global_norm_clip_factor(gd, 0.5)
        if kind == "SGD":
            ops.momentum_step(p, gd, s1, lr, 0.9, 1.0, fac)
        else:
            ops.rmsprop_step(p, gd, s1, s2, lr, 0.9, 0.0, 1e-10, 1.0, fac)
    assert rel(p, params["p"]) < 1e-5


def test_clip_factor_uses_the_mean_gradient_norm(ops):
    """under data parallelism the buffer holds the SUM over G ranks; tf.clip_by_global_norm
    (models/network.py:191-192) acts on the batch-mean gradient, so the norm must be taken of grad_scale * g."""
    g = torch.Generator().manual_seed(5)
    mean_grad = torch.randn(5000, generator=g)
    G, clip = 4, 3.0
    summed = dev(mean_grad * G)
    fac = float(ops.global_norm_clip_factor(summed, clip, 1.0 / G))
    want = clip / max(float(mean_grad.double().norm()), clip)
    assert abs(fac - want) < 1e-5 * want
    assert abs(float(ops.global_norm_clip_factor(summed, clip)) - clip / float((mean_grad * G).double().norm())) < 1e-6


def test_prepare_inputs_mix_and_normalize(ops):
    """Input contract on the device (data/dataset.py:456-468): mixture = sequential fp32 sum of the sources (bit-exact vs
    numpy's sum over the stacked axis), --dataset_normalize = per-source (x - mean) / sqrt(var)."""
    rng = np.random.RandomState(3)
    for S, Lw in ((2, 4096), (3, 1001)):
        nm = (rng.randn(3, S, Lw) * 0.05 + 0.01).astype(np.float32)
        x_mix, _ = ops.prepare_inputs(dev(nm))
        assert np.array_equal(x_mix.cpu().numpy(), nm.sum(1))
        d = dev(nm.copy())
        x_mix, stats = ops.prepare_inputs(d, normalize=True)
        t = torch.tensor(nm)
        mean, var = t.mean(-1, keepdim=True), t.var(-1, unbiased=False, keepdim=True)
        want = (t - mean) / torch.sqrt(var)
        assert rel(d, want) < 1e-5
        assert rel(stats[:, 0], mean.reshape(-1)) < 1e-4 and rel(stats[:, 1], var.reshape(-1)) < 1e-4
        assert rel(x_mix, want.sum(1)) < 1e-5


# ------------------------------------------------------------------------------------------ fused Adapt front-output terms
@pytest.mark.parametrize("S,separation", [(2, "mask"), (2, "perfect"), (3, "perfect")])
def test_adapt_terms_fwd_bwd_match_oracle(ops, S, separation):
    """amss_adapt_terms_fwd / _bwd (models/adapt.py:127-132, 141-160, 162-196, 315-316): separator output, p_hat, the three
    scalar terms and d(weighted sum)/dy against torch autograd of the oracle's restatement."""
    g = torch.Generator().manual_seed(60 + S)
    B, Tp, N = 3, 9, 16
    y = (torch.randn(B * (S + 1), Tp, N, generator=g, dtype=torch.float64) * 0.02).requires_grad_(True)
    rho = 0.01
    p_hat = y.abs().reshape(y.shape[0], -1).sum(0)
    sparse = T.kl_div(rho, p_hat).sum()
    overlap = M.adapt_overlap(y, B, S)
    sep = M.adapt_separator_pretraining(y, B, S, separation)
    neg = (torch.where(y < 0, y, torch.zeros_like(y)) ** 2).reshape(y.shape[0], -1).sum(1).mean()
    gsep = torch.randn(sep.shape, generator=g, dtype=torch.float64)
    w = torch.tensor([0.7, -1.3, 2.1], dtype=torch.float64)
    total = (sep * gsep).sum() + w[0] * sparse + w[1] * overlap + w[2] * neg
    (dy_ref,) = torch.autograd.grad(total, y)
    yd = dev(y.detach().float())
    sep_d, ph_d, terms = ops.adapt_terms_fwd(yd, B, S, rho, 0 if separation == "mask" else 1)
    assert rel(sep_d, sep) < 1e-5 and rel(ph_d, p_hat) < 1e-5
    for got, want in zip(terms.cpu().tolist(), (sparse, overlap, neg)):
        assert abs(got - float(want)) < 1e-4 * max(abs(float(want)), 1e-6), (got, float(want))
    dy = ops.adapt_terms_bwd(yd, ph_d, dev(gsep.float()), dev(w.float()), B, S, rho, 0 if separation == "mask" else 1)
    assert rel(dy, dy_ref) < 1e-4


@pytest.mark.parametrize("S,nl", [(2, "softmax"), (3, "softmax"), (2, "tanh"), (2, "None")])
def test_enhance_cost_fused_matches_oracle(S, nl):
    """amss_enhance_cost_table / _bwd (models/network.py:640-693): cost and d cost / d logits vs torch autograd of the oracle."""
    import amss_b200  # noqa: F401
    from amss_b200 import layers
    g = torch.Generator().manual_seed(90 + S)
    B, TF = 3, 777
    logits = torch.randn(B, S, TF, generator=g, dtype=torch.float64).requires_grad_(True)
    X = torch.rand(B, TF, generator=g, dtype=torch.float64) + 0.1
    tgt = torch.rand(B, TF, S, generator=g, dtype=torch.float64)
    yv = logits.transpose(1, 2)
    yv = torch.softmax(yv, -1) if nl == "softmax" else (torch.tanh(yv) if nl == "tanh" else yv)
    cost = M.enhance_cost(yv * X.unsqueeze(-1), tgt)
    (gref,) = torch.autograd.grad(cost, logits)
    ld = dev(logits.detach().float()).requires_grad_(True)
    c = layers.enhance_cost_fused(ld, dev(X.float()), dev(tgt.float()), nl)
    (gd,) = torch.autograd.grad(c, ld)
    assert abs(float(c) - float(cost)) < 1e-4 * abs(float(cost))
    assert rel(gd, gref) < 1e-4
