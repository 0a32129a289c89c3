"""amss_b200 -- B200-native (sm_100a) implementation of the data-parallel hot path of
Totoketchup/Adaptive-MultiSpeaker-Separation: adaptive conv filterbank / STFT twin -> stacked
BLSTM embeddings (DPCL / L41) -> k-means masks -> waveform inversion, fwd + bwd + AMSGrad.

Importing the package loads libamss_b200.so (hand-written CUDA behind the C ABI of
include/amss.h) and raises if it is missing: there is no CPU or PyTorch fallback.
The directory name carries a hyphen, so import it through the `amss_b200` shim at the repo root.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is absent)
from ._lib import AmssError, AMSS_PREC_FP32, AMSS_PREC_BF16, AMSS_POOL_MAX, AMSS_POOL_AVG, AMSS_POOL_STRIDE  # noqa: F401

__all__ = ["AmssError", "AMSS_PREC_FP32", "AMSS_PREC_BF16", "AMSS_POOL_MAX", "AMSS_POOL_AVG", "AMSS_POOL_STRIDE"]
