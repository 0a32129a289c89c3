"""GPU tests of the tcgen05 (bf16 operands, fp32 accumulate) kernels, through the C ABI.

The parity gate proper (1e-3 relative, north_star) is evaluated on the fp32 kernels in
test_gpu_kernels.py.  bf16 operands carry 8 significant bits, so these tests compare the
tensor-core kernels with the oracle evaluated ON bf16-ROUNDED OPERANDS (fp32 accumulate, which is
what the hardware computes): tolerance 2e-3 relative to the largest reference value, written in
each test; index outputs (arg-max) must point at a position whose reference value is within that
tolerance of the true maximum (near-ties may legitimately flip)."""
import numpy as np
import pytest
import torch

from oracle import tf_ops as T

pytestmark = pytest.mark.gpu

TOL_BF16 = 2e-3


@pytest.fixture(scope="module")
def ops():
    import amss_b200  # noqa: F401
    from amss_b200 import ops as o
    return o


def dev(x):
    return torch.as_tensor(x).cuda().contiguous()


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


# (Bt, L, W, N, pool)
ANALYSIS_CASES = [
    (3, 4096, 64, 16, 32),        # smoke-sized: partial filter tile, several frames per time tile
    (2, 8192, 1024, 256, 256),    # the bench geometry, short signal
    (2, 5000, 200, 130, 128),     # ragged: W not a multiple of 64, N spills into a 2nd m-tile, L % pool != 0
    (1, 4096, 96, 128, 512),      # frames spanning two time tiles
    (2, 4096, 1024, 512, 512),    # the reference's DEFAULT bank (utils/trainer.py:136-147): four filter tiles, pool 512
]


@pytest.mark.parametrize("Bt,L,W,N,pool", ANALYSIS_CASES)
def test_analysis_tc_matches_oracle_on_bf16_operands(ops, Bt, L, W, N, pool):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(Bt, L, generator=g) * 0.05
    filt = torch.randn(W, N, generator=g) / np.sqrt(W)
    xr, fr = bf16_round(x), bf16_round(filt)
    X = T.conv2d_same_1d(xr, fr)                          # [Bt, L, N] fp32 on the rounded operands
    Tp = (L - pool) // pool + 1
    ref_y, ref_am = T.max_pool_with_argmax_1d(X, pool, pool)
    y, am = ops.filterbank_analysis(dev(x), dev(filt), pool, pool, ops.AMSS_POOL_MAX, ops.AMSS_PREC_BF16)
    assert y.shape == (Bt, Tp, N) and am.shape == (Bt, Tp, N)
    scale = float(ref_y.abs().max())
    assert float((y.cpu() - ref_y).abs().max()) <= TOL_BF16 * scale
    # arg-max: TF flat index t*N + n; the chosen position must hold a (near-)maximal reference value
    am = am.cpu()
    n_idx = am % N
    t_idx = am // N
    assert torch.equal(n_idx, torch.arange(N).expand(Bt, Tp, N))
    frame = torch.arange(Tp).view(1, Tp, 1)
    assert bool(((t_idx >= frame * pool) & (t_idx < (frame + 1) * pool)).all())
    picked = torch.gather(X, 1, t_idx)                      # X[b, t_idx[b,tp,n], n]
    assert float((picked - ref_y).abs().max()) <= TOL_BF16 * scale
    agree = float((am == ref_am).float().mean())
    assert agree > 0.98, agree


@pytest.mark.parametrize("B,L,W,N,pool", [(2, 4096, 64, 16, 128), (3, 8192, 1024, 256, 256), (1, 5000, 200, 130, 256),
                                          (2, 4096, 1024, 512, 512)])
def test_analysis_mix_linear_mixture_path(ops, B, L, W, N, pool):
    """amss_filterbank_analysis_mix_fwd, S = 2: when x_mix == x_0 + x_1 bit for bit the mixture rows come from the sum of
    the two source responses.  Source rows must equal the stock kernel's bit for bit (same products, same order);
    mixture rows must match conv(bf16(x_0)) + conv(bf16(x_1)) pooled, and the fp32 kernel within the bf16 tolerance.
    When the equality does not hold (one sample perturbed) the call must return exactly what the stock kernel returns."""
    g = torch.Generator().manual_seed(13)
    src = torch.randn(B, 2, L, generator=g) * 0.05
    mix = src[:, 0] + src[:, 1]
    x = torch.cat([mix, src.reshape(B * 2, L)], 0).contiguous()
    filt = torch.randn(W, N, generator=g) / np.sqrt(W)
    P = ops.AMSS_PREC_BF16
    y_stock, am_stock = ops.filterbank_analysis(dev(x), dev(filt), pool, pool, ops.AMSS_POOL_MAX, P)
    y, am = ops.filterbank_analysis_mix(dev(x), dev(filt), B, 2, pool, pool, P)
    assert torch.equal(y[B:], y_stock[B:]) and torch.equal(am[B:], am_stock[B:])          # source rows: identical
    fr = bf16_round(filt)
    Xs = T.conv2d_same_1d(bf16_round(src.reshape(B * 2, L)), fr).reshape(B, 2, L, N)
    ref_y, ref_am = T.max_pool_with_argmax_1d(Xs[:, 0] + Xs[:, 1], pool, pool)
    scale = float(ref_y.abs().max())
    assert float((y[:B].cpu() - ref_y).abs().max()) <= TOL_BF16 * scale
    assert float((am[:B].cpu() == ref_am).float().mean()) > 0.98
    assert float((y[:B] - y_stock[:B]).abs().max()) <= 1e-2 * scale                      # vs conv(bf16(x_mix))
    assert not torch.equal(y[:B], y_stock[:B])                                            # (the fast path really ran)
    # broken contract: the device-side check must route the batch through the stock kernel
    x2 = x.clone()
    x2[0, L // 2] += 1e-3
    y2, am2 = ops.filterbank_analysis_mix(dev(x2), dev(filt), B, 2, pool, pool, P)
    y2s, am2s = ops.filterbank_analysis(dev(x2), dev(filt), pool, pool, ops.AMSS_POOL_MAX, P)
    assert torch.equal(y2, y2s) and torch.equal(am2, am2s)
    # three sources per mixture: stock path
    x3 = torch.randn(4, L, generator=g) * 0.05
    y3, _ = ops.filterbank_analysis_mix(dev(x3), dev(filt), 1, 3, pool, pool, P)
    y3s, _ = ops.filterbank_analysis(dev(x3), dev(filt), pool, pool, ops.AMSS_POOL_MAX, P)
    assert torch.equal(y3, y3s)


def test_analysis_tc_close_to_fp32_kernel(ops):
    """bf16-operand result vs the fp32 SIMT kernel on the same inputs: reports the operand-rounding
    error (about 2^-9 per product, averaged over W=1024 taps) and bounds it at 1e-2 of the peak."""
    g = torch.Generator().manual_seed(12)
    x = torch.randn(3, 16384, generator=g) * 0.05
    filt = torch.randn(1024, 256, generator=g) / 32.0
    y32, am32 = ops.filterbank_analysis(dev(x), dev(filt), 256, 256, ops.AMSS_POOL_MAX, ops.AMSS_PREC_FP32)
    y16, am16 = ops.filterbank_analysis(dev(x), dev(filt), 256, 256, ops.AMSS_POOL_MAX, ops.AMSS_PREC_BF16)
    err = float((y16 - y32).abs().max() / y32.abs().max())
    agree = float((am16 == am32).float().mean())
    print(f"bf16 vs fp32 analysis: max rel err {err:.2e}, arg-max agreement {agree:.4f}")
    assert err < 1e-2
    assert agree > 0.9


# ------------------------------------------------------------------------------------------ GEMM
def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("shape", [(128, 256, 64), (37, 53, 29), (300, 1200, 257), (1, 7, 3), (260, 520, 1000),
                                   (600, 1200, 4000)])
def test_gemm_tc_matches_bf16_operand_product(ops, ta, tb, shape):
    Mm, N, K = shape
    g = torch.Generator().manual_seed(21)
    A = torch.randn((K, Mm) if ta else (Mm, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    Ar, Br = bf16_round(A), bf16_round(B)
    ref = (Ar.t() if ta else Ar).double() @ (Br.t() if tb else Br).double() + bias.double()
    out = ops.gemm(dev(A), dev(B), dev(bias), bool(ta), bool(tb), precision=ops.AMSS_PREC_BF16)
    assert rel(out, ref) < 1e-4          # fp32 accumulation of exact bf16 products


@pytest.mark.parametrize("M,N,K,E", [(300, 1280, 600, 40), (77, 96, 50, 16), (129, 480, 64, 48), (260, 10240, 600, 40)])
def test_gemm_bf16_fused_l2_normalize_epilogue(ops, M, N, K, E):
    """amss_gemm_bf16 with norm_E: normalised rows and inv_norm equal l2_normalize(x W + b) over groups of E columns
    (tf.nn.l2_normalize, utils/ops.py:323-324) computed from the same bf16 operands; one all-zero group exercises the
    clamped branch (inv_norm stored negative, as amss_l2norm_fwd does)."""
    g = torch.Generator().manual_seed(23)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(K, N, generator=g) * 0.1
    b = torch.randn(N, generator=g) * 0.1
    W[:, E:2 * E] = 0.0
    b[E:2 * E] = 0.0                                                   # group 1 is exactly zero
    xb, Wb = ops.convert_bf16(dev(x)), ops.convert_bf16(dev(W))
    assert xb.dtype == torch.bfloat16 and xb.shape[1] % 8 == 0
    V, inv = ops.gemm_bf16(xb, False, Wb, True, M, N, K, bias=dev(b), norm_E=E)
    z = bf16_round(x).double() @ bf16_round(W).double() + b.double()
    zg = z.view(M, N // E, E)
    ss = (zg * zg).sum(-1, keepdim=True)
    ref = (zg / ss.clamp_min(1e-12).sqrt()).view(M, N)
    assert rel(V, ref) < 1e-4
    inv_ref = (1.0 / ss.clamp_min(1e-12).sqrt()).view(M, N // E)
    inv_h = inv.view(M, N // E).double().cpu()
    assert bool((inv_h[:, 1] < 0).all())                               # clamped group flagged
    keep = torch.ones(N // E, dtype=torch.bool)
    keep[1] = False
    assert float(((inv_h - inv_ref)[:, keep].abs() / inv_ref[:, keep]).max()) < 1e-4
    # plain bf16 GEMM through the same entry point, both operand majors
    out = ops.gemm_bf16(xb, False, Wb, True, M, N, K, bias=dev(b))
    assert rel(out, z) < 1e-4
    dzb = ops.convert_bf16(dev(torch.randn(M, N, generator=g)))
    dW = ops.gemm_bf16(xb, True, dzb, True, K, N, M)
    assert rel(dW, bf16_round(x).double().t() @ dzb.double().cpu()[:, :N]) < 1e-4
    dx = ops.gemm_bf16(dzb, False, Wb, False, M, K, N)
    assert rel(dx, dzb.double().cpu()[:, :N] @ bf16_round(W).double().t()) < 1e-4


@pytest.mark.parametrize("D0,D1,C", [(7, 5, 600), (13, 3, 37), (250, 4, 256)])
def test_transpose_01_bf16_is_transpose_then_round(ops, D0, D1, C):
    g = torch.Generator().manual_seed(24)
    x = torch.randn(D0, D1, C, generator=g)
    out = ops.transpose_01_bf16(dev(x))
    Cp = (C + 7) // 8 * 8
    assert out.shape == (D1 * D0, Cp) and out.dtype == torch.bfloat16
    ref = x.transpose(0, 1).reshape(D1 * D0, C).to(torch.bfloat16)
    assert torch.equal(out[:, :C].cpu(), ref)
    assert bool((out[:, C:] == 0).all())


def test_gemm_tc_strided_accumulate_swap(ops):
    g = torch.Generator().manual_seed(22)
    Tt, Bb, C, N = 50, 6, 40, 72
    big = torch.randn(Tt * Bb, 2 * C, generator=g)
    A = dev(big)[:, C:]                       # row-strided view (lda = 2C)
    W = torch.randn(C, N, generator=g)
    base = torch.randn(Tt * Bb, N, generator=g)
    out = dev(base.clone())
    ops.gemm(A, dev(W), None, out=out, accumulate=True, precision=ops.AMSS_PREC_BF16)
    prod = bf16_round(big[:, C:]).double() @ bf16_round(W).double()
    assert rel(out, base.double() + prod) < 1e-4
    out2 = ops.gemm(A, dev(W), None, out_swap=(Bb, Tt), precision=ops.AMSS_PREC_BF16)
    ref2 = prod.reshape(Tt, Bb, N).transpose(0, 1).reshape(Bb * Tt, N)
    assert rel(out2, ref2) < 1e-4
    # split-K with accumulate (weight-gradient shape): C += A^T B
    X = torch.randn(4000, 96, generator=g)
    dZ = torch.randn(4000, 200, generator=g)
    acc0 = torch.randn(96, 200, generator=g)
    out3 = dev(acc0.clone())
    ops.gemm(dev(X), dev(dZ), None, transa=True, out=out3, accumulate=True, precision=ops.AMSS_PREC_BF16)
    ref3 = acc0.double() + bf16_round(X).double().t() @ bf16_round(dZ).double()
    assert rel(out3, ref3) < 1e-4


# ------------------------------------------------------------------------------------------ BLSTM
def _saved_views(saved, B, Tt, H):
    """amss_blstm_saved_bytes layout: gates [2][T][B][4H] (256-byte aligned), then cst [2][T][B][H]."""
    f = saved.view(torch.float32)
    ng = 2 * Tt * B * 4 * H
    off = (ng * 4 + 255) // 256 * 256 // 4
    return f[:ng], f[off:off + 2 * Tt * B * H]


BLSTM_CASES = [(3, 7, 11, 6), (5, 20, 40, 150), (40, 12, 64, 300), (70, 9, 33, 20), (16, 30, 256, 300)]


@pytest.mark.parametrize("B,Tt,I,H", BLSTM_CASES)
def test_blstm_tc_fwd_bwd_close_to_oracle(ops, B, Tt, I, H):
    """bf16 weights / hidden state / dz on the tensor cores, MUFU tanh, bf16 partial-sum exchange:
    forward within 1e-2 of the output range, gradients within 3e-2 of each gradient's range."""
    import math
    g = torch.Generator().manual_seed(40 + B)
    x = (torch.randn(B, Tt, I, generator=g, dtype=torch.float64) * 0.5).requires_grad_(True)
    ks = [((torch.rand(I + H, 4 * H, generator=g, dtype=torch.float64) * 2 - 1) * math.sqrt(6.0 / (I + 5 * H))).requires_grad_(True)
          for _ in range(2)]
    bs = [(torch.randn(4 * H, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True) for _ in range(2)]
    dy = torch.randn(B, Tt, 2 * H, generator=g, dtype=torch.float64)
    y = T.blstm(x, ks[0], bs[0], ks[1], bs[1])
    grads = torch.autograd.grad((y * dy).sum(), [x] + ks + bs)
    f = lambda t: dev(t.detach().float())
    x_tm = ops.transpose_01(f(x))
    args = (x_tm, f(ks[0]), f(bs[0]), f(ks[1]), f(bs[1]))
    y_tm, saved = ops.blstm_fwd(*args, precision=ops.AMSS_PREC_BF16)
    y32, saved32 = ops.blstm_fwd(*args, precision=ops.AMSS_PREC_FP32)
    err = rel(ops.transpose_01(y_tm), y)
    g16, c16 = _saved_views(saved, B, Tt, H)
    g32, c32 = _saved_views(saved32, B, Tt, H)
    eg, ec = rel(g16, g32), rel(c16, c32)
    dy_tm = ops.transpose_01(f(dy))
    dx, dkf, dbf, dkb, dbb = ops.blstm_bwd(x_tm, f(ks[0]), f(ks[1]), y_tm, dy_tm, saved, precision=ops.AMSS_PREC_BF16)
    errs = [rel(ops.transpose_01(dx), grads[0]), rel(dkf, grads[1]), rel(dkb, grads[2]), rel(dbf, grads[3]),
            rel(dbb, grads[4])]
    print(f"blstm tc B={B} T={Tt} I={I} H={H}: y {err:.2e} gates {eg:.2e} c {ec:.2e} "
          f"dx {errs[0]:.2e} dK {errs[1]:.2e}/{errs[2]:.2e} db {errs[3]:.2e}/{errs[4]:.2e}")
    assert err < 1e-2 and eg < 2e-2 and ec < 2e-2
    assert max(errs) < 3e-2


# ------------------------------------------------------------------------------------------ whole steps in bf16
def test_training_steps_bf16_track_the_oracle():
    """Whole fwd+bwd+AMSGrad steps with every tensor-core kernel switched on (precision='bf16'): the cost of each
    step stays within 2 % of the fp32 oracle's and the parameter update points the same way (cosine > 0.98)."""
    import functools
    from amss_b200 import models, trainer
    from oracle import models as M
    from oracle import steps as OS
    B, S, Lw = 3, 2, 4096
    t = trainer.Front_Separator_Trainer(models.DPCL, nb_layers=2, layer_size=300, embedding_size=8, learning_rate=1e-3,
                                        window_size=64, filters=16, max_pool=32, hop_size=32, with_max_pool=True,
                                        precision="bf16")
    p = {k: v.detach().cpu().clone() for k, v in t.store.params.items()}
    p0 = {k: v.clone() for k, v in p.items()}
    st = OS.Stepper(p, functools.partial(OS.front_separator_loss, nb_layers=2, embedding_size=8, max_pool=32, hop=32), lr=1e-3)
    for step in range(2):
        mix, nm, I = M.synthetic_mixtures(B, S, Lw, seed=500 + step)
        c_ref, _ = st.step(torch.tensor(mix), torch.tensor(nm), torch.tensor(I))
        c = float(t.train_step(dev(mix), dev(nm), dev(I)))
        assert abs(c - c_ref) < 2e-2 * abs(c_ref), (step, c, c_ref)
    num = den_a = den_b = 0.0
    for k, v in st.tr.items():
        da = (t.store[k].detach().cpu() - p0[k]).double().reshape(-1)
        db = (v.detach() - p0[k]).double().reshape(-1)
        num += float(da @ db); den_a += float(da @ da); den_b += float(db @ db)
    cos = num / (den_a ** 0.5 * den_b ** 0.5 + 1e-30)
    print(f"bf16 step: update cosine vs fp32 oracle {cos:.4f}")
    assert cos > 0.98


@pytest.mark.parametrize("E", [10, 56])
def test_stft_dpcl_config1_width_bf16_step(E):
    """BASELINE config 1 widths (2 x BLSTM-300 -> H = 150 per direction, NC = 5 CTAs per cluster) in bf16.  E = 10: not a
    multiple of 4, the loss runs on the fp32 DPCL kernels behind the bf16 trunk.  E = 56: wider than the fused-normalise
    epilogue handles (48), so the head takes the unfused dense -> l2_normalize path and the loss node still moves onto
    the head's (x, W, b) with the bf16 dz hand-over (_amss_dense side channel)."""
    import functools
    from amss_b200 import models, trainer
    from oracle import models as M
    from oracle import steps as OS
    B, S, Lw = 4, 2, 4096
    t = trainer.STFT_Separator_Trainer(models.DPCL, nb_layers=2, layer_size=300, embedding_size=E, learning_rate=1e-3,
                                       window_size=128, hop_size=64, precision="bf16")
    p = {k: v.detach().cpu().clone() for k, v in t.store.params.items()}
    st = OS.Stepper(p, functools.partial(OS.stft_separator_loss, nb_layers=2, embedding_size=E, window_size=128, hop_size=64),
                    lr=1e-3)
    mix, nm, I = M.synthetic_mixtures(B, S, Lw, seed=600)
    c_ref, _ = st.step(torch.tensor(mix), torch.tensor(nm), torch.tensor(I))
    c = float(t.train_step(dev(mix), dev(nm), dev(I)))
    assert abs(c - c_ref) < 2e-2 * abs(c_ref), (c, c_ref)


def test_l41_bf16_step_uses_the_generic_head_backward():
    """L41 (sigmoid-dot loss on the normalised embeddings) in bf16: the fused dense + l2-normalise forward with the
    GENERIC backward of _DenseNormFn (dV -> dz -> dx, dW, db), i.e. the path taken when the DPCL cost does not follow."""
    import functools
    from amss_b200 import models, trainer
    from oracle import models as M
    from oracle import steps as OS
    B, S, Lw = 3, 2, 4096
    t = trainer.STFT_Separator_Trainer(models.L41Model, nb_layers=1, layer_size=64, embedding_size=8, learning_rate=1e-3,
                                       window_size=128, hop_size=64, precision="bf16")
    p = {k: v.detach().cpu().clone() for k, v in t.store.params.items()}
    p0 = {k: v.clone() for k, v in p.items()}
    st = OS.Stepper(p, functools.partial(OS.stft_separator_loss, nb_layers=1, embedding_size=8, window_size=128, hop_size=64,
                                         loss="l41"), lr=1e-3)
    mix, nm, I = M.synthetic_mixtures(B, S, Lw, seed=700)
    c_ref, _ = st.step(torch.tensor(mix), torch.tensor(nm), torch.tensor(I))
    c = float(t.train_step(dev(mix), dev(nm), dev(I)))
    assert abs(c - c_ref) < 2e-2 * abs(c_ref), (c, c_ref)
    num = den_a = den_b = 0.0
    for k, v in st.tr.items():
        da = (t.store[k].detach().cpu() - p0[k]).double().reshape(-1)
        db = (v.detach() - p0[k]).double().reshape(-1)
        num += float(da @ db); den_a += float(da @ da); den_b += float(db @ db)
    assert num / (den_a ** 0.5 * den_b ** 0.5 + 1e-30) > 0.97


def test_dpcl_backward_bf16_output_matches_fp32_output(ops):
    """amss_dpcl_loss_bwd_normalized_bf16 writes the same dz as the fp32-output tensor-core variant, rounded to bf16, and
    amss_colsum_bf16 sums it like amss_colsum sums the fp32 one."""
    g = torch.Generator().manual_seed(91)
    B, TF, E, S = 3, 2000, 40, 2
    z = dev(torch.randn(B, TF, E, generator=g))
    lab = dev(torch.randint(0, S, (B, TF), generator=g).to(torch.uint8))
    V, inv = ops.l2norm_fwd(z, E)
    loss, ws = ops.dpcl_loss_fwd(V, lab, S, ops.AMSS_PREC_BF16)
    one = torch.ones(1, device="cuda")
    dz32 = ops.dpcl_loss_bwd_normalized(V, lab, S, one, ws, inv, ops.AMSS_PREC_BF16)
    dzb = ops.dpcl_loss_bwd_normalized_bf16(V, lab, S, one, ws, inv)
    assert dzb.dtype == torch.bfloat16
    assert torch.equal(dzb, dz32.to(torch.bfloat16))
    cs = ops.colsum_bf16(dzb.view(B * TF // 50, 50 * E))
    ref = dzb.view(B * TF // 50, 50 * E).double().sum(0)
    assert rel(cs, ref) < 1e-5


# ------------------------------------------------------------------------------------------ fused DPCL backward
@pytest.mark.parametrize("B,TF,E,S", [(3, 1000, 40, 2), (2, 130, 40, 3), (5, 128, 16, 2), (2, 4099, 64, 2), (70, 257, 40, 2)])
def test_dpcl_fused_backward_tc_matches_autograd(ops, B, TF, E, S):
    """dz of  loss(l2_normalize(z))  from the fused tcgen05 kernel vs torch autograd of the oracle (fp64) and vs the fp32
    fused kernel: 5e-3 of the gradient range (bf16 operands in the [points x E] x [E x E] product)."""
    from oracle import models as M
    g = torch.Generator().manual_seed(70 + B)
    lab = torch.randint(0, S, (B, TF), generator=g)
    # clustered (anisotropic) embeddings with imperfect labels, so that An v is far from parallel to v and the
    # tensor-core product really matters in dz
    centers = torch.randn(B, S, E, generator=g, dtype=torch.float64) * 2.0
    noisy = torch.where(torch.rand(B, TF, generator=g) < 0.3, torch.randint(0, S, (B, TF), generator=g), lab)
    z = centers[torch.arange(B)[:, None], noisy] + 0.5 * torch.randn(B, TF, E, generator=g, dtype=torch.float64)
    z[0, 3] = 0.0                                          # a row in the clamped branch of l2_normalize
    z.requires_grad_(True)
    V = T.l2_normalize(z, -1)
    cost = M.dpcl_cost(V.reshape(B, TF, 1, E), torch.nn.functional.one_hot(lab, S).double().reshape(B, TF, 1, S))
    (gz,) = torch.autograd.grad(cost, z)
    zf = dev(z.detach().float())
    labd = dev(lab.to(torch.uint8))
    Vg, inv = ops.l2norm_fwd(zf, E)
    loss, ws = ops.dpcl_loss_fwd(Vg, labd, S)
    loss16, ws16 = ops.dpcl_loss_fwd(Vg, labd, S, ops.AMSS_PREC_BF16)      # Gram statistics on tcgen05
    assert abs(float(loss) - float(cost)) < 1e-3 * abs(float(cost))
    assert abs(float(loss16) - float(cost)) < 3e-3 * abs(float(cost)), (float(loss16), float(cost))
    one = torch.ones(1, device="cuda")
    dz32 = ops.dpcl_loss_bwd_normalized(Vg, labd, S, one, ws, inv, ops.AMSS_PREC_FP32)
    dz16 = ops.dpcl_loss_bwd_normalized(Vg, labd, S, one, ws16, inv, ops.AMSS_PREC_BF16)
    # the clamped row has inv_norm = 1e6 and would dominate a max-norm comparison: check it on its own
    keep = torch.ones(B, TF, dtype=torch.bool)
    keep[0, 3] = False
    gk = gz[keep]
    assert rel(dz32.cpu()[keep], gk) < 1e-3 and rel(dz32.cpu()[0, 3], gz[0, 3]) < 1e-3
    err = rel(dz16.cpu()[keep], gk)
    print(f"fused dpcl bwd tc B={B} TF={TF} E={E} S={S}: rel err {err:.2e} (fp32 kernel {rel(dz32.cpu()[keep], gk):.1e})")
    assert 1e-6 < err < 5e-3                               # bf16 operands: visibly not the fp32 kernel, within tolerance
    assert rel(dz16.cpu()[0, 3], gz[0, 3]) < 5e-3
