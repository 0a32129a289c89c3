"""Full-size (BASELINE.json geometry: L = 64000 samples, W = 1024 taps, 256 filters, pool 256, T = 250, 3 x BLSTM-600,
E = 40) checks through size-independent properties (the comparisons with the ORACLE at this geometry are in
test_gpu_fullsize_oracle.py): each test states a
property of the reference operator that must hold at any size, plus agreement between the fp32 parity kernels and the
tensor-core kernels on the same inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

L, W, N, POOL = 64000, 1024, 256, 256


@pytest.fixture(scope="module")
def ops():
    import amss_b200  # noqa: F401
    from amss_b200 import ops as o
    return o


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def test_analysis_full_size_homogeneity_and_tc_agreement(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(3, L, device="cuda", generator=g) * 0.05
    filt = torch.randn(W, N, device="cuda", generator=g) / 32
    y32, am32 = ops.filterbank_analysis(x, filt, POOL, POOL, ops.AMSS_POOL_MAX, ops.AMSS_PREC_FP32)
    assert y32.shape == (3, 250, N)
    # max-pooling of a linear map is positively homogeneous: scale the input by a power of two -> same arg-max, scaled max
    y2, am2 = ops.filterbank_analysis(4.0 * x, filt, POOL, POOL, ops.AMSS_POOL_MAX, ops.AMSS_PREC_FP32)
    assert torch.equal(am2, am32) and torch.equal(y2, 4.0 * y32)
    # the arg-max is a valid per-sample flat index t*N + n inside its pooling window
    t_idx, n_idx = am32 // N, am32 % N
    frame = torch.arange(250, device="cuda").view(1, 250, 1)
    assert bool(((t_idx >= frame * POOL) & (t_idx < (frame + 1) * POOL)).all())
    assert torch.equal(n_idx, torch.arange(N, device="cuda").expand_as(n_idx))
    # tensor-core kernel on the same inputs (bf16 operands): values within 1e-2 of the peak, arg-max mostly identical
    y16, am16 = ops.filterbank_analysis(x, filt, POOL, POOL, ops.AMSS_POOL_MAX, ops.AMSS_PREC_BF16)
    assert rel(y16, y32) < 1e-2
    assert float((am16 == am32).float().mean()) > 0.9
    y16b, am16b = ops.filterbank_analysis(4.0 * x, filt, POOL, POOL, ops.AMSS_POOL_MAX, ops.AMSS_PREC_BF16)
    assert torch.equal(am16b, am16) and torch.equal(y16b, 4.0 * y16)          # exact: power-of-two scaling commutes with bf16


def test_synthesis_is_the_adjoint_of_unpooled_analysis_full_size(ops):
    """<synthesis(v), z> == <v, d synthesis / d v applied to z> at L = 64000 (linear operator and its transpose)."""
    g = torch.Generator(device="cuda").manual_seed(2)
    B, S = 2, 2
    Tp = (L - POOL) // POOL + 1
    x = torch.randn(B, L, device="cuda", generator=g) * 0.05
    filt = torch.randn(W, N, device="cuda", generator=g) / 32
    _, am = ops.filterbank_analysis(x, filt, POOL, POOL, ops.AMSS_POOL_MAX, ops.AMSS_PREC_FP32)
    v = torch.randn(B * S, Tp, N, device="cuda", generator=g)
    z = torch.randn(B * S, L, device="cuda", generator=g)
    out = ops.filterbank_synthesis(v, am, filt, B, S, L, POOL, POOL)
    dv, _ = ops.filterbank_synthesis_bwd(z, v, am, filt, B, S, need_dvals=True, need_dfilt=False)
    lhs, rhs = float((out.double() * z.double()).sum()), float((v.double() * dv.double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), abs(rhs), 1.0)


def test_istft_inverts_stft_full_size(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, L, device="cuda", generator=g) * 0.05
    spec, mag = ops.stft(x, 512, 256)
    assert spec.shape == (2, 249, 257)
    ones = torch.zeros(2, 249 * 257, dtype=torch.int32, device="cuda")
    rec = ops.istft_masked(spec, 1, 512, 256, labels=ones)[:, 0]
    n = rec.shape[1]
    assert rel(rec[:, 256:n - 256], x[:, 256:n - 256]) < 1e-4          # the edges are attenuated by TF's window normalisation


def test_kmeans_assignments_are_a_fixed_point_full_size(ops):
    """K = 3 (config 5) on TF = 63993 points: re-running the assignment from the returned centroids reproduces the labels."""
    g = torch.Generator(device="cuda").manual_seed(4)
    B, TF, E, K = 2, 63993, 40, 3
    centers = torch.randn(B, K, E, device="cuda", generator=g) * 3
    assign = torch.randint(0, K, (B, TF), device="cuda", generator=g)
    X = centers[torch.arange(B, device="cuda")[:, None], assign] + 0.05 * torch.randn(B, TF, E, device="cuda", generator=g)
    init = torch.stack([torch.randperm(TF, generator=torch.Generator().manual_seed(i))[:K] for i in range(B * 4)]).to(torch.int32).cuda()
    cent, labels, inertia, best = ops.kmeans_fit(X, init, K, 4, 10)
    Xn = torch.nn.functional.normalize(X, dim=-1)
    d = ((Xn[:, :, None, :] - cent[:, None, :, :]) ** 2).sum(-1)
    assert torch.equal(d.argmin(-1).to(torch.int32), labels)
    # well separated blobs: the partition equals the generating one up to a permutation of the cluster ids
    for b in range(B):
        m = torch.zeros(K, K)
        for i in range(K):
            for j in range(K):
                m[i, j] = ((labels[b] == i) & (assign[b] == j)).sum()
        assert int((m > 0).sum()) == K


def test_blstm_full_size_tc_tracks_fp32(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    T, B, I, H = 250, 4, 256, 300
    x = torch.randn(T, B, I, device="cuda", generator=g) * 0.3
    k = [(torch.rand(I + H, 4 * H, device="cuda", generator=g) * 2 - 1) * (6.0 / (I + 5 * H)) ** 0.5 for _ in range(2)]
    b = [torch.zeros(4 * H, device="cuda") for _ in range(2)]
    y32, s32 = ops.blstm_fwd(x, k[0], b[0], k[1], b[1], precision=ops.AMSS_PREC_FP32)
    y16, s16 = ops.blstm_fwd(x, k[0], b[0], k[1], b[1], precision=ops.AMSS_PREC_BF16)
    assert rel(y16, y32) < 1e-2
    dy = torch.randn_like(y32)
    g32 = ops.blstm_bwd(x, k[0], k[1], y32, dy, s32, precision=ops.AMSS_PREC_FP32)
    g16 = ops.blstm_bwd(x, k[0], k[1], y16, dy, s16, precision=ops.AMSS_PREC_BF16)
    for a, c in zip(g16, g32):
        assert rel(a, c) < 3e-2


def test_bench_step_full_size_fp32_vs_bf16(ops):
    """One full-size training step of the bench workload (B = 2): same loss in fp32 and bf16, finite gradients, the loss
    goes down over a few AMSGrad steps."""
    from amss_b200 import models, trainer, synth
    cfg = dict(nb_speakers=2, nb_layers=3, layer_size=600, embedding_size=40, window_size=1024, filters=256, max_pool=256,
               hop_size=256, with_max_pool=True, learning_rate=1e-3)
    mix, nm, I = synth.synthetic_mixtures(2, 2, L, seed=9)
    batch = [torch.as_tensor(a).cuda() for a in (mix, nm, I)]
    t32 = trainer.Front_Separator_Trainer(models.DPCL, precision="fp32", **cfg)
    t16 = trainer.Front_Separator_Trainer(models.DPCL, precision="bf16", **cfg)
    t16.store.load_state_dict(t32.store.state_dict())
    c32 = float(t32.train_step(*batch))
    losses = [float(t16.train_step(*batch)) for _ in range(4)]
    assert abs(losses[0] - c32) < 2e-2 * abs(c32), (losses[0], c32)
    assert all(np.isfinite(losses)) and bool(torch.isfinite(t16.store.grad_flat).all())
    assert losses[-1] < losses[0]


def test_linear_mixture_front_end_full_size(ops):
    """BASELINE geometry, 4 mixtures of 2 sources: the linear-mixture path of amss_filterbank_analysis_mix_fwd must (a)
    leave the source rows bit-identical to the stock tensor-core kernel, (b) produce mixture rows that agree with the
    fp32 kernel run on x_mix within the bf16 tolerance and pick an arg-max whose fp32 response is within that tolerance
    of the true maximum, and (c) be positively homogeneous like the operator it replaces."""
    g = torch.Generator(device="cuda").manual_seed(6)
    B = 4
    src = torch.randn(B, 2, L, device="cuda", generator=g) * 0.05
    x = torch.cat([src[:, 0] + src[:, 1], src.reshape(2 * B, L)], 0).contiguous()
    filt = torch.randn(W, N, device="cuda", generator=g) / 32
    y, am = ops.filterbank_analysis_mix(x, filt, B, 2, POOL, POOL, ops.AMSS_PREC_BF16)
    ys, ams = ops.filterbank_analysis(x, filt, POOL, POOL, ops.AMSS_POOL_MAX, ops.AMSS_PREC_BF16)
    assert torch.equal(y[B:], ys[B:]) and torch.equal(am[B:], ams[B:])
    assert not torch.equal(y[:B], ys[:B])                                   # the mixture rows came from the linear path
    y32, am32 = ops.filterbank_analysis(x[:B].contiguous(), filt, POOL, POOL, ops.AMSS_POOL_MAX, ops.AMSS_PREC_FP32)
    scale = float(y32.abs().max())
    assert float((y[:B] - y32).abs().max()) < 1e-2 * scale
    assert float((am[:B] == am32).float().mean()) > 0.9
    y4, am4 = ops.filterbank_analysis_mix(4.0 * x, filt, B, 2, POOL, POOL, ops.AMSS_PREC_BF16)
    assert torch.equal(am4, am) and torch.equal(y4, 4.0 * y)


def test_head_gemm_fused_normalise_full_size(ops):
    """[T*B, 600] x [600, 10240] head product with the l2-normalise epilogue at the bench geometry (8 mixtures): every
    group of E = 40 output columns has unit norm, inv_norm reproduces the norm of the plain product, and
    V * norm equals the plain bf16 product."""
    g = torch.Generator(device="cuda").manual_seed(7)
    M, K, NN, E = 250 * 8, 600, 10240, 40
    h = torch.randn(M, K, device="cuda", generator=g)
    Wt = torch.randn(K, NN, device="cuda", generator=g) * 0.05
    b = torch.randn(NN, device="cuda", generator=g) * 0.1
    hb, Wb = ops.convert_bf16(h), ops.convert_bf16(Wt)
    V, inv = ops.gemm_bf16(hb, False, Wb, True, M, NN, K, bias=b, norm_E=E)
    z = ops.gemm_bf16(hb, False, Wb, True, M, NN, K, bias=b)
    nrm = V.view(M, NN // E, E).double().pow(2).sum(-1).sqrt()
    assert float((nrm - 1.0).abs().max()) < 1e-5
    zn = z.view(M, NN // E, E).double().pow(2).sum(-1).sqrt()
    assert float((inv.view(M, NN // E).double() * zn - 1.0).abs().max()) < 1e-4
    assert rel(V.view(M, NN // E, E) * zn.unsqueeze(-1).float(), z.view(M, NN // E, E)) < 1e-5
