"""The oracle has no reference golden vectors to be pinned against (parity unpinned,
see oracle/__init__.py); every TF-op restatement is therefore cross-checked here against
an independent formulation."""
import math

import numpy as np
import pytest
import scipy.signal
import torch

from oracle import tf_ops as T
from oracle import models as M
from oracle.kmeans import KMeans, random_init_idx
from oracle.amsgrad import AMSGrad

torch.manual_seed(0)


def test_conv_same_matches_definition():
    x = torch.randn(2, 50, dtype=torch.float64)
    f = torch.randn(8, 3, dtype=torch.float64)
    X = T.conv2d_same_1d(x, f, 1)
    W = 8
    pl = (W - 1) // 2
    ref = torch.zeros(2, 50, 3, dtype=torch.float64)
    for t in range(50):
        for k in range(W):
            s = t + k - pl
            if 0 <= s < 50:
                ref[:, t] += x[:, s:s + 1] * f[k]
    assert torch.allclose(X, ref, atol=1e-12)
    assert T.same_pad_1d(64000, 1024, 1) == (64000, 511, 512)


def test_strided_conv_same_padding():
    x = torch.randn(1, 37, dtype=torch.float64)
    f = torch.randn(6, 2, dtype=torch.float64)
    s = 4
    out, pl, pr = T.same_pad_1d(37, 6, s)
    assert out == 10 and pl + pr == max((out - 1) * s + 6 - 37, 0)
    X = T.conv2d_same_1d(x, f, s)
    assert X.shape == (1, 10, 2)
    ref = torch.zeros(1, 10, 2, dtype=torch.float64)
    for t in range(10):
        for k in range(6):
            i = t * s + k - pl
            if 0 <= i < 37:
                ref[:, t] += x[:, i:i + 1] * f[k]
    assert torch.allclose(X, ref, atol=1e-12)


def test_transpose_is_adjoint_of_conv():
    L, W, N = 64, 10, 4
    x = torch.randn(3, L, dtype=torch.float64)
    U = torch.randn(3, L, N, dtype=torch.float64)
    f = torch.randn(W, N, dtype=torch.float64)
    lhs = (T.conv2d_same_1d(x, f, 1) * U).sum()
    rhs = (x * T.conv2d_transpose_same_1d(U, f, L, 1)).sum()
    assert abs(lhs - rhs) < 1e-9


def test_maxpool_argmax_and_unpool_roundtrip():
    X = torch.randn(2, 40, 3)
    X[0, 5, 1] = X[0, 6, 1] = 100.0   # tie inside window 0 -> first wins
    y, am = T.max_pool_with_argmax_1d(X, 8, 8)
    assert y.shape == (2, 5, 3) and am.dtype == torch.int64
    assert am[0, 0, 1].item() == 5 * 3 + 1
    U = T.unpool(y, am, 40, 3)
    assert torch.equal(U.reshape(2, -1).gather(1, am.reshape(2, -1)).reshape(y.shape), y)
    assert (U != 0).sum() == y.numel()
    # overlapping windows (ksize > stride)
    y2, am2 = T.max_pool_with_argmax_1d(X, 8, 4)
    assert y2.shape[1] == (40 - 8) // 4 + 1
    t = am2 // 3
    assert torch.all((t >= (torch.arange(y2.shape[1]) * 4).view(1, -1, 1)) & (t < (torch.arange(y2.shape[1]) * 4 + 8).view(1, -1, 1)))


def test_stft_matches_scipy_and_istft_reconstructs():
    L = 4096
    x = torch.randn(2, L, dtype=torch.float64)
    S = T.stft(x, 512, 256)
    assert S.shape == (2, 1 + (L - 512) // 256, 257)
    f, t, Z = scipy.signal.stft(x.numpy(), window=scipy.signal.get_window("hann", 512, fftbins=True),
                                nperseg=512, noverlap=256, boundary=None, padded=False)
    Z = Z * scipy.signal.get_window("hann", 512, fftbins=True).sum()   # scipy scales by 1/sum(w)
    assert np.allclose(S.numpy(), np.transpose(Z, (0, 2, 1)), atol=1e-9)
    rec = T.inverse_stft(S, 512, 256)
    assert rec.shape[1] == L
    assert torch.allclose(rec[:, 256:L - 256], x[:, 256:L - 256], atol=1e-10)
    # edges are attenuated, not reconstructed (TF window normalisation)
    assert not torch.allclose(rec[:, :256], x[:, :256], atol=1e-3)


def test_basic_lstm_matches_torch_lstm():
    B, Tt, I, H = 3, 7, 5, 4
    x = torch.randn(B, Tt, I, dtype=torch.float64)
    kf, kb = torch.randn(I + H, 4 * H, dtype=torch.float64) * 0.3, torch.randn(I + H, 4 * H, dtype=torch.float64) * 0.3
    bf, bb = torch.randn(4 * H, dtype=torch.float64) * 0.1, torch.randn(4 * H, dtype=torch.float64) * 0.1
    out = T.blstm(x, kf, bf, kb, bb)
    lstm = torch.nn.LSTM(I, H, batch_first=True, bidirectional=True).double()

    def load(kernel, bias, sfx):
        i, j, f, o = kernel.split(H, 1)
        bi, bj, bff, bo = bias.split(H)
        Wt = torch.cat([i, f, j, o], 1)                      # torch order i,f,g,o
        bt = torch.cat([bi, bff + 1.0, bj, bo])              # forget_bias = 1.0
        getattr(lstm, "weight_ih_l0" + sfx).data = Wt[:I].t().contiguous()
        getattr(lstm, "weight_hh_l0" + sfx).data = Wt[I:].t().contiguous()
        getattr(lstm, "bias_ih_l0" + sfx).data = bt
        getattr(lstm, "bias_hh_l0" + sfx).data = torch.zeros_like(bt)

    load(kf, bf, "")
    load(kb, bb, "_reverse")
    ref, _ = lstm(x)
    assert torch.allclose(out, ref, atol=1e-10)


def test_dpcl_cost_matches_closed_form():
    B, Tt, Fb, S, E = 2, 6, 5, 2, 4
    V = T.l2_normalize(torch.randn(B, Tt, Fb, E, dtype=torch.float64), 3)
    lab = torch.randint(0, S, (B, Tt, Fb))
    lab[:, 0, 0] = 0
    lab[:, 0, 1] = 1
    y = torch.nn.functional.one_hot(lab, S).double()
    c = M.dpcl_cost(V, y)
    ref = 0.0
    for b in range(B):
        Vb = V[b].reshape(-1, E)
        Yb = y[b].reshape(-1, S)
        cnt = Yb.sum(0)
        d = 1.0 / torch.sqrt(Yb @ cnt)
        A = Vb.t() @ (d[:, None] * Vb)
        C = Vb.t() @ (d[:, None] * Yb)
        ref += torch.linalg.norm(A) - 2 * torch.linalg.norm(C) + torch.sqrt((torch.sqrt(cnt) ** 2).sum())
    assert abs(c - ref / B) < 1e-10


def test_kmeans_recovers_blobs():
    from sklearn.datasets import make_blobs
    X, ytrue = make_blobs(n_samples=1000, centers=4, n_features=40, cluster_std=1.0, random_state=3)
    Xb = np.stack([X, X, X]).astype(np.float32)
    km = KMeans(4, nb_tries=10, nb_iterations=10, normalize_input=False)
    rng = np.random.RandomState(0)
    cent, labels = km.fit(Xb, random_init_idx(30, 1000, 4, rng))
    labels = labels.numpy()
    for b in range(3):
        # partition equality up to label permutation
        pairs = set(zip(labels[b].tolist(), ytrue.tolist()))
        assert len(pairs) == 4


def test_kmeans_silence_quirk_and_soft():
    torch.manual_seed(1)
    X = torch.randn(2, 200, 8)
    lat = torch.rand(2, 200) + 1e-3
    lat[:, :50] = 1e-6                                    # silent bins
    rng = np.random.RandomState(1)
    idx = random_init_idx(2 * 3, 200, 2, rng)
    idx = np.where(idx < 50, idx + 60, idx)               # init on non-silent rows
    km = KMeans(2, nb_tries=3, nb_iterations=4, threshold=2.0, assign_at_end=False)
    _, labels = km.fit(X, idx, latent=lat)
    assert labels.dtype == torch.int32 and torch.all(labels[:, :50] == 0)
    kms = KMeans(2, nb_tries=3, nb_iterations=4, beta=5.0)
    _, soft = kms.fit(X, idx)
    assert soft.shape == (2, 200, 2) and torch.allclose(soft.sum(-1), torch.ones(2, 200), atol=1e-5)


def test_amsgrad_matches_manual():
    p = {"w": torch.tensor([1.0, -2.0, 3.0])}
    opt = AMSGrad(p, lr=0.1)
    g1 = torch.tensor([0.5, -1.0, 2.0])
    w0 = p["w"].clone()
    opt.step({"w": g1})
    m = 0.1 * g1
    v = 0.01 * g1 * g1
    lr_t = 0.1 * math.sqrt(1 - 0.99) / (1 - 0.9)
    assert torch.allclose(p["w"], w0 - lr_t * m / (torch.sqrt(v) + 1e-3), atol=1e-7)
    opt.step({"w": torch.zeros(3)})                       # vhat keeps the max
    assert torch.allclose(opt.vhat["w"], v)


def test_adapt_pretraining_mask_autoencoder_shapes_and_grads():
    torch.manual_seed(0)
    B, S, L, W, N, P = 2, 2, 256, 32, 8, 16
    p = {k: v.clone().requires_grad_(True) for k, v in M.init_adapt_params(W, N, dtype=torch.float64).items()}
    xnm = torch.randn(B, S, L, dtype=torch.float64) * 0.1
    xm = xnm.sum(1)
    cost, aux = M.adapt_pretraining_cost(p, xm, xnm, max_pool=P, hop=P, loss="sdr+l2", separation="mask",
                                         beta=0.01, overlap_coef=1e-3)
    assert aux["y"].shape == (B * (S + 1), (L - P) // P + 1, N)
    assert aux["back"].shape == (B, S, L)
    cost.backward()
    assert all(torch.isfinite(v.grad).all() for v in p.values())
    # 'perfect' separation feeds mix - (sum others) = own representation only when fronts are linear in time positions
    cost2, _ = M.adapt_pretraining_cost(p, xm, xnm, max_pool=P, hop=P, loss="l2", separation="perfect")
    assert torch.isfinite(cost2)


def test_separator_stft_pipeline_and_l41():
    torch.manual_seed(0)
    xm, xnm, I = M.synthetic_mixtures(2, 2, 4096, seed=1)
    xm, xnm, I = torch.tensor(xm), torch.tensor(xnm), torch.tensor(I)
    assert torch.allclose(xm, xnm.sum(1), atol=1e-6)
    pre = M.separator_preprocessing(xm, xnm, 512, 256, 1.0, -1.0)
    assert pre["y"].shape == (2, 15, 257, 2)
    assert set(pre["y"].unique().tolist()) == {-1.0, 1.0}
    p = M.init_separator_params(257, 1, 8, 5, with_speaker_vectors=True)
    V = M.separator_prediction(p, pre["X"], 1, 5)
    assert torch.allclose((V ** 2).sum(-1), torch.ones(2, 15, 257), atol=1e-5)
    c = M.l41_cost(p, V, pre["y"], I)
    assert torch.isfinite(c) and c > 0
    # separate + postprocessing with oracle masks reproduces each source's dominant bins
    km = KMeans(2, nb_tries=2, nb_iterations=3)
    rng = np.random.RandomState(0)
    fn = lambda emb: km.fit(emb, random_init_idx(emb.shape[0] * 2, emb.shape[1], 2, rng))[1]
    sep, masks = M.separate(V, pre["X"], fn, 2)
    out = M.postprocessing(sep, pre["stfts"], 2, 512, 256)
    assert out.shape == (2, 2, 4096)
    # masks partition the mixture: summed estimates reconstruct the mixture interior
    assert torch.allclose(out.sum(1)[:, 256:-256], xm[:, 256:-256], atol=1e-4)


def test_cost_finetuning_is_permutation_invariant_pit():
    """cost_finetuning (models/network.py:697-723): 0.5*sum_L (x - xhat)^2, mean over S, MIN over the S! permutations,
    mean over B -- zero for any permutation of the targets, invariant to permuting the estimates, equal to the
    identity-permutation l2 when the estimates are already aligned, and its gradient pulls every estimate towards the
    target it is matched with."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 3, 400, generator=g, dtype=torch.float64)
    for perm in ([0, 1, 2], [2, 0, 1], [1, 0, 2]):
        assert float(M.cost_finetuning(x, x[:, perm])) == 0.0
    est = (x + 0.05 * torch.randn(3, 3, 400, generator=g, dtype=torch.float64)).requires_grad_(True)
    c = M.cost_finetuning(x, est)
    direct = (0.5 * ((x - est) ** 2).sum(-1)).mean(-1).mean()
    assert abs(float(c) - float(direct)) < 1e-12
    assert abs(float(M.cost_finetuning(x, est[:, [1, 2, 0]])) - float(c)) < 1e-12
    grad, = torch.autograd.grad(c, est)
    assert torch.allclose(grad, (est - x).detach() / 9.0, atol=1e-12)          # 1/(B*S) * (xhat - x)


def test_momentum_and_rmsprop_and_decay():
    """--optimizer SGD / RMSProp (models/network.py:175-186): Momentum against torch.optim.SGD (same recurrence),
    RMSProp against the closed form of TF's kernel (rms slot starts at ONE, epsilon inside the square root) and the
    staircase decay lr * 0.5^(epoch // decay_epoch)."""
    from oracle.amsgrad import Momentum, RMSProp, exponential_decay
    g = torch.Generator().manual_seed(0)
    w0 = torch.randn(7, generator=g)
    grads = [torch.randn(7, generator=g) for _ in range(4)]
    p = {"w": w0.clone()}
    opt = Momentum(p, lr=0.1, decay_epoch=2)
    wt = w0.clone().requires_grad_(True)
    sgd = torch.optim.SGD([wt], lr=0.1, momentum=0.9)
    for gr in grads[:2]:
        opt.step({"w": gr})
        wt.grad = gr.clone()
        sgd.step()
    assert torch.allclose(p["w"], wt.detach(), atol=1e-6)
    opt.increment_epoch(); opt.increment_epoch()               # epoch 2 -> lr halves (decay_epoch = 2)
    before = p["w"].clone()
    acc = opt.accum["w"].clone()
    opt.step({"w": grads[2]})
    assert torch.allclose(p["w"], before - 0.05 * (0.9 * acc + grads[2]), atol=1e-7)
    assert exponential_decay(0.1, 49, 50) == 0.1 and exponential_decay(0.1, 50, 50) == 0.05 and exponential_decay(0.1, 149, 50) == 0.025
    q = {"w": w0.clone()}
    rms = RMSProp(q, lr=0.01)
    ms = torch.ones(7)
    w = w0.clone()
    for gr in grads[:3]:
        rms.step({"w": gr})
        ms = 0.9 * ms + 0.1 * gr * gr
        w = w - 0.01 * gr / torch.sqrt(ms + 1e-10)
    assert torch.allclose(q["w"], w, atol=1e-7)
