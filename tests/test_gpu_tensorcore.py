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
