"""Functional host wrappers over the C ABI (include/amss.h).

torch is used for device memory and the current stream only; every computation below is a
kernel of libamss_b200.so.  All functions take/return CUDA float32 tensors (contiguous) unless
noted, and raise if handed CPU tensors -- there is no CPU path.
"""
import math

import torch

from . import _lib
from ._lib import AMSS_PREC_FP32, AMSS_PREC_BF16, AMSS_POOL_MAX, AMSS_POOL_AVG, AMSS_POOL_STRIDE  # noqa: F401

_f32 = torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.AmssError("amss ops need CUDA tensors: there is no CPU fallback")
        if not t.is_contiguous():
            raise _lib.AmssError("amss ops need contiguous tensors")


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------
# adaptive filterbank
# ------------------------------------------------------------------------------------------
def make_filter(window, bases):
    _chk(window, bases)
    W, N = bases.shape
    filt = torch.empty(W, N, dtype=_f32, device=bases.device)
    _lib.call("amss_filterbank_make_filter", _p(window), _p(bases), W, N, _p(filt), _stream())
    return filt


def make_filter_bwd(window, bases, dfilt):
    _chk(window, bases, dfilt)
    W, N = bases.shape
    dwindow = torch.empty_like(window)
    dbases = torch.empty_like(bases)
    _lib.call("amss_filterbank_make_filter_bwd", _p(window), _p(bases), _p(dfilt), W, N, _p(dwindow), _p(dbases),
              _stream())
    return dwindow, dbases


def analysis_out_frames(L, W, pool, hop, mode):
    return _lib.query("amss_filterbank_analysis_out_frames", L, W, pool, hop, mode)


def filterbank_analysis(x, filt, pool, hop, mode=AMSS_POOL_MAX, precision=AMSS_PREC_FP32):
    """x[Bt,L], filt[W,N] -> (y[Bt,Tp,N], argmax int64 [Bt,Tp,N] or None)."""
    _chk(x, filt)
    Bt, L = x.shape
    W, N = filt.shape
    Tp = analysis_out_frames(L, W, pool, hop, mode)
    y = torch.empty(Bt, Tp, N, dtype=_f32, device=x.device)
    am = torch.empty(Bt, Tp, N, dtype=torch.int64, device=x.device) if mode == AMSS_POOL_MAX else None
    nb = _lib.query("amss_filterbank_analysis_workspace_bytes", Bt, L, W, N, pool, hop, mode, precision)
    ws = _ws(nb, x.device)
    _lib.call("amss_filterbank_analysis_fwd", _p(x), _p(filt), Bt, L, W, N, pool, hop, mode, precision, _p(y), _p(am),
              _p(ws), ws.numel(), _stream())
    return y, am


def filterbank_analysis_mix(x, filt, B, S, pool, hop, precision=AMSS_PREC_FP32):
    """Front end of a training batch x = [B mixtures ; B*S sources] (max-pool mode): same outputs as
    filterbank_analysis; the library derives the mixture rows by linearity when x_mix == sum of its sources bit for bit
    (checked on the device per call, see amss_filterbank_analysis_mix_fwd)."""
    _chk(x, filt)
    Bt, L = x.shape
    if Bt != B * (S + 1):
        raise _lib.AmssError(f"filterbank_analysis_mix: {Bt} rows for B={B}, S={S}")
    W, N = filt.shape
    Tp = analysis_out_frames(L, W, pool, hop, AMSS_POOL_MAX)
    y = torch.empty(Bt, Tp, N, dtype=_f32, device=x.device)
    am = torch.empty(Bt, Tp, N, dtype=torch.int64, device=x.device)
    nb = _lib.query("amss_filterbank_analysis_workspace_bytes", Bt, L, W, N, pool, hop, AMSS_POOL_MAX, precision)
    ws = _ws(nb, x.device)
    _lib.call("amss_filterbank_analysis_mix_fwd", _p(x), _p(filt), B, S, L, W, N, pool, hop, AMSS_POOL_MAX, precision, _p(y),
              _p(am), _p(ws), ws.numel(), _stream())
    return y, am


def filterbank_analysis_bwd(x, dy, argmax, W):
    """d(filt)[W,N] through the max-pool arg-max."""
    _chk(x, dy, argmax)
    Bt, L = x.shape
    _, Tp, N = dy.shape
    dfilt = torch.empty(W, N, dtype=_f32, device=x.device)
    ws = _ws(_lib.query("amss_filterbank_grad_workspace_bytes", W, N), x.device)
    _lib.call("amss_filterbank_analysis_bwd", _p(x), _p(dy), _p(argmax), Bt, L, W, N, Tp, 0, _p(dfilt), _p(ws),
              ws.numel(), _stream())
    return dfilt


def box_sum(x, P, dir=1, scale=1.0, len_out=None, axis=-1):
    """Sliding box sum along `axis` of a contiguous 2-D tensor: out[.., u, ..] = scale * sum_{j<P} x[.., u + dir*j, ..]
    (zero outside).  axis=-1: rows are the series; axis=0: columns are the series ([W, N] filter banks)."""
    _chk(x)
    assert x.dim() == 2 and x.is_contiguous()
    if axis in (-1, 1):
        nser, len_in = x.shape
        lo = len_in if len_out is None else int(len_out)
        out = torch.empty(nser, lo, dtype=_f32, device=x.device)
        _lib.call("amss_box_sum", _p(x), nser, len_in, len_in, 1, int(P), int(dir), float(scale), lo, lo, 1, _p(out), _stream())
    else:
        len_in, nser = x.shape
        lo = len_in if len_out is None else int(len_out)
        out = torch.empty(lo, nser, dtype=_f32, device=x.device)
        _lib.call("amss_box_sum", _p(x), nser, len_in, 1, nser, int(P), int(dir), float(scale), lo, 1, nser, _p(out), _stream())
    return out


def filterbank_synthesis(vals, argmax_mix, filt2, B, S, L, pool, hop):
    """vals[B*S,Tp,N], argmax_mix[B,Tp,N] (mixture rows), filt2[W,N] -> out[B*S,L]."""
    _chk(vals, argmax_mix, filt2)
    R, Tp, N = vals.shape
    W = filt2.shape[0]
    out = torch.empty(R, L, dtype=_f32, device=vals.device)
    ws = _ws(_lib.query("amss_filterbank_synthesis_workspace_bytes", B, S, L, W, N, Tp), vals.device)
    _lib.call("amss_filterbank_synthesis_fwd", _p(vals), _p(argmax_mix), _p(filt2), B, S, L, W, N, Tp, pool, hop,
              _p(out), _p(ws), ws.numel(), _stream())
    return out


def filterbank_synthesis_bwd(dout, vals, argmax_mix, filt2, B, S, need_dvals=True, need_dfilt=True):
    _chk(dout, vals, argmax_mix, filt2)
    R, Tp, N = vals.shape
    W = filt2.shape[0]
    L = dout.shape[1]
    dvals = torch.empty_like(vals) if need_dvals else None
    dfilt2 = torch.empty_like(filt2) if need_dfilt else None
    ws = _ws(_lib.query("amss_filterbank_grad_workspace_bytes", W, N), vals.device)
    _lib.call("amss_filterbank_synthesis_bwd", _p(dout), _p(vals), _p(argmax_mix), _p(filt2), B, S, L, W, N, Tp,
              _p(dvals), _p(dfilt2), _p(ws), ws.numel(), _stream())
    return dvals, dfilt2


# ------------------------------------------------------------------------------------------
# STFT twin
# ------------------------------------------------------------------------------------------
def stft(x, frame, hop, want_spec=True, want_mag=True):
    """x[R,L] -> (spec complex64 [R,T,F] or None, mag [R,T,F] or None)."""
    _chk(x)
    R, L = x.shape
    T, F = 1 + (L - frame) // hop, frame // 2 + 1
    spec = torch.empty(R, T, F, 2, dtype=_f32, device=x.device) if want_spec else None
    mag = torch.empty(R, T, F, dtype=_f32, device=x.device) if want_mag else None
    _lib.call("amss_stft_fwd", _p(x), R, L, frame, hop, _p(spec), _p(mag), _stream())
    return (torch.view_as_complex(spec) if want_spec else None), mag


def stft_labels(non_mix, frame, hop, want_mag=False):
    """non_mix[B,S,L] -> (labels uint8 [B,T,F], |X_non_mix| [B,T,F,S] or None)."""
    _chk(non_mix)
    B, S, L = non_mix.shape
    T, F = 1 + (L - frame) // hop, frame // 2 + 1
    labels = torch.empty(B, T, F, dtype=torch.uint8, device=non_mix.device)
    mag = torch.empty(B, T, F, S, dtype=_f32, device=non_mix.device) if want_mag else None
    _lib.call("amss_stft_labels", _p(non_mix), B, S, L, frame, hop, _p(labels), _p(mag), _stream())
    return labels, mag


def istft_masked(spec, S, frame, hop, labels=None, masks=None):
    """spec complex64 [B,T,F]; labels int32 [B,T*F] or masks [B,T*F,S] -> out[B,S,L']."""
    specr = torch.view_as_real(spec).contiguous()
    _chk(specr, labels, masks)
    B, T, F = spec.shape
    Lout = (T - 1) * hop + frame
    out = torch.empty(B, S, Lout, dtype=_f32, device=spec.device)
    _lib.call("amss_istft_masked_fwd", _p(specr), _p(labels), _p(masks), B, S, T, frame, hop, _p(out), _stream())
    return out


def istft_masked_bwd(spec, dout, S, frame, hop):
    """Gradient of istft_masked(spec, masks=...) w.r.t. the soft masks: dout[B,S,L'] -> dmasks[B,T*F,S]."""
    specr = torch.view_as_real(spec).contiguous()
    _chk(specr, dout)
    B, T, F = spec.shape
    dmasks = torch.empty(B, T * F, S, dtype=_f32, device=spec.device)
    _lib.call("amss_istft_masked_bwd", _p(specr), _p(dout), B, S, T, frame, hop, _p(dmasks), _stream())
    return dmasks


# ------------------------------------------------------------------------------------------
# GEMM / BLSTM
# ------------------------------------------------------------------------------------------
def transpose_01(x):
    """x[D0,D1,C] -> [D1,D0,C]."""
    _chk(x)
    D0, D1, C = x.shape
    out = torch.empty(D1, D0, C, dtype=_f32, device=x.device)
    _lib.call("amss_transpose_01", _p(x), D0, D1, C, _p(out), _stream())
    return out


def gemm(A, B, bias=None, transa=False, transb=False, out=None, accumulate=False, precision=AMSS_PREC_FP32,
         out_swap=None):
    """op(A) @ op(B) (+bias).  A, B 2-D, possibly row-strided views (stride(1) == 1).
    out_swap=(b_count, t_count) remaps time-major output rows to batch-major."""
    for t in (A, B):
        if not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
            raise _lib.AmssError("gemm: operands must be 2-D CUDA tensors with unit inner stride")
    M = A.shape[1] if transa else A.shape[0]
    K = A.shape[0] if transa else A.shape[1]
    Kb = B.shape[1] if transb else B.shape[0]
    N = B.shape[0] if transb else B.shape[1]
    if K != Kb:
        raise _lib.AmssError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, dtype=_f32, device=A.device)
        accumulate = False
    sb, st = (out_swap if out_swap else (0, 0))
    nb = _lib.query("amss_gemm_workspace_bytes", M, N, K, int(transa), int(transb), precision)
    ws = _ws(nb, A.device)
    _lib.call("amss_gemm", _p(A), A.stride(0), _p(B), B.stride(0), _p(bias), M, N, K, int(transa), int(transb),
              int(accumulate), precision, _p(out), out.stride(0), sb, st, _p(ws), ws.numel(), _stream())
    return out


def transpose_01_bf16(x):
    """x[D0,D1,C] fp32 -> bf16 [D1*D0, pad8(C)] (rows in [D1][D0] order, zero padded): transpose_01 + convert_bf16 fused."""
    _chk(x)
    D0, D1, C = x.shape
    ldd = (C + 7) // 8 * 8
    out = torch.empty(D1 * D0, ldd, dtype=torch.bfloat16, device=x.device)
    _lib.call("amss_transpose_01_bf16", _p(x), D0, D1, C, _p(out), ldd, _stream())
    return out


def convert_bf16(x):
    """fp32 [rows, cols] (unit inner stride) -> bf16 [rows, pad8(cols)] (zero padded): operand of gemm_bf16."""
    if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1:
        raise _lib.AmssError("convert_bf16: need a 2-D CUDA tensor with unit inner stride")
    rows, cols = x.shape
    ldd = (cols + 7) // 8 * 8
    out = torch.empty(rows, ldd, dtype=torch.bfloat16, device=x.device)
    _lib.call("amss_convert_bf16", _p(x), rows, cols, x.stride(0), _p(out), ldd, _stream())
    return out


def gemm_bf16(A, a_mn, B, b_mn, M, N, K, bias=None, out=None, accumulate=False, out_swap=None, norm_E=0):
    """C[M,N] (+)= A B from bf16 operands (see amss_gemm_bf16): a_mn False -> A [M,K]; True -> A stored [K,M];
    b_mn True -> B [K,N]; False -> B stored [N,K].  norm_E > 0: returns (l2-normalised C, inv_norm[M*N/norm_E])."""
    for t in (A, B):
        if not t.is_cuda or t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.bfloat16:
            raise _lib.AmssError("gemm_bf16: operands must be 2-D bf16 CUDA tensors with unit inner stride")
    if out is None:
        out = torch.empty(M, N, dtype=_f32, device=A.device)
        accumulate = False
    sb, st = (out_swap if out_swap else (0, 0))
    inv = torch.empty(M * (N // norm_E), dtype=_f32, device=A.device) if norm_E else None
    _lib.call("amss_gemm_bf16", _p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(bias), M, N, K,
              int(accumulate), _p(out), out.stride(0), sb, st, int(norm_E), _p(inv), _stream())
    return (out, inv) if norm_E else out


def blstm_fwd(x_tm, kernel_fw, bias_fw, kernel_bw, bias_bw, forget_bias=1.0, precision=AMSS_PREC_FP32,
              save_for_backward=True):
    """x_tm[T,B,I] time-major -> (y_tm[T,B,2H], saved or None)."""
    _chk(x_tm, kernel_fw, bias_fw, kernel_bw, bias_bw)
    T, B, I = x_tm.shape
    H = kernel_fw.shape[1] // 4
    assert kernel_fw.shape[0] == I + H and kernel_bw.shape == kernel_fw.shape
    y = torch.empty(T, B, 2 * H, dtype=_f32, device=x_tm.device)
    saved = _ws(_lib.query("amss_blstm_saved_bytes", B, T, I, H), x_tm.device) if save_for_backward else None
    ws = _ws(_lib.query("amss_blstm_workspace_bytes", B, T, I, H, precision), x_tm.device)
    _lib.call("amss_blstm_fwd", _p(x_tm), _p(kernel_fw), _p(bias_fw), _p(kernel_bw), _p(bias_bw), B, T, I, H,
              float(forget_bias), precision, _p(y), _p(saved), _p(ws), ws.numel(), _stream())
    return y, saved


def blstm_bwd(x_tm, kernel_fw, kernel_bw, y_tm, dy_tm, saved, precision=AMSS_PREC_FP32, need_dx=True):
    _chk(x_tm, kernel_fw, kernel_bw, y_tm, dy_tm, saved)
    T, B, I = x_tm.shape
    H = kernel_fw.shape[1] // 4
    dx = torch.empty_like(x_tm) if need_dx else None
    dk_fw, dk_bw = torch.empty_like(kernel_fw), torch.empty_like(kernel_bw)
    db_fw = torch.empty(4 * H, dtype=_f32, device=x_tm.device)
    db_bw = torch.empty(4 * H, dtype=_f32, device=x_tm.device)
    ws = _ws(_lib.query("amss_blstm_workspace_bytes", B, T, I, H, precision), x_tm.device)
    _lib.call("amss_blstm_bwd", _p(x_tm), _p(kernel_fw), _p(kernel_bw), _p(y_tm), _p(dy_tm), _p(saved), B, T, I, H,
              precision, _p(dx), _p(dk_fw), _p(db_fw), _p(dk_bw), _p(db_bw), _p(ws), ws.numel(), _stream())
    return dx, dk_fw, db_fw, dk_bw, db_bw


# ------------------------------------------------------------------------------------------
# normalisation / losses
# ------------------------------------------------------------------------------------------
def l2norm_fwd(z, E):
    _chk(z)
    rows = z.numel() // E
    v = torch.empty_like(z)
    inv = torch.empty(rows, dtype=_f32, device=z.device)
    _lib.call("amss_l2norm_fwd", _p(z), rows, E, _p(v), _p(inv), _stream())
    return v, inv


def l2norm_bwd(v, inv, dv, E):
    _chk(v, inv, dv)
    dz = torch.empty_like(v)
    _lib.call("amss_l2norm_bwd", _p(v), _p(inv), _p(dv), v.numel() // E, E, _p(dz), _stream())
    return dz


def colsum(dZ):
    _chk(dZ)
    M, N = dZ.shape
    out = torch.empty(N, dtype=_f32, device=dZ.device)
    ws = _ws(_lib.query("amss_colsum_workspace_bytes", M, N), dZ.device)
    _lib.call("amss_colsum", _p(dZ), M, N, _p(out), _p(ws), ws.numel(), _stream())
    return out


def colsum_bf16(dZ):
    _chk(dZ)
    M, N = dZ.shape
    out = torch.empty(N, dtype=_f32, device=dZ.device)
    ws = _ws(_lib.query("amss_colsum_workspace_bytes", M, N), dZ.device)
    _lib.call("amss_colsum_bf16", _p(dZ), M, N, _p(out), _p(ws), ws.numel(), _stream())
    return out


def dpcl_loss_fwd(V, labels, S, precision=AMSS_PREC_FP32):
    """V[B,TF,E], labels uint8 [B,TF] -> (loss[1], workspace for the backward)."""
    _chk(V, labels)
    B, TF, E = V.shape
    loss = torch.empty(1, dtype=_f32, device=V.device)
    ws = _ws(_lib.query("amss_dpcl_workspace_bytes", B, TF, E, S), V.device)
    _lib.call("amss_dpcl_loss_fwd_prec", _p(V), _p(labels), B, TF, E, S, precision, _p(loss), _p(ws), ws.numel(), _stream())
    return loss, ws


def dpcl_loss_weighted_fwd(V, labels, weights, S):
    """DPCL cost with Y = weights * one_hot(labels) (--function_mask): -> (loss[1], workspace for the backward)."""
    _chk(V, labels, weights)
    B, TF, E = V.shape
    loss = torch.empty(1, dtype=_f32, device=V.device)
    ws = _ws(_lib.query("amss_dpcl_workspace_bytes", B, TF, E, S), V.device)
    _lib.call("amss_dpcl_loss_weighted_fwd", _p(V), _p(labels), _p(weights), B, TF, E, S, _p(loss), _p(ws), ws.numel(),
              _stream())
    return loss, ws


def dpcl_loss_weighted_bwd(V, labels, weights, S, dloss, ws, inv_norm=None):
    """dV, or dz through the l2_normalize Jacobian when inv_norm is given."""
    _chk(V, labels, weights, dloss)
    B, TF, E = V.shape
    out = torch.empty_like(V)
    _lib.call("amss_dpcl_loss_weighted_bwd", _p(V), _p(labels), _p(weights), _p(dloss),
              _p(inv_norm), B, TF, E, S, _p(out), _p(ws), _stream())
    return out


def dpcl_loss_bwd(V, labels, S, dloss, ws):
    _chk(V, labels, dloss)
    B, TF, E = V.shape
    dV = torch.empty_like(V)
    _lib.call("amss_dpcl_loss_bwd", _p(V), _p(labels), _p(dloss), B, TF, E, S, _p(dV), _p(ws), _stream())
    return dV


def dpcl_loss_bwd_normalized(V, labels, S, dloss, ws, inv_norm, precision=AMSS_PREC_FP32):
    """Fused DPCL backward + l2_normalize backward: returns dz (gradient w.r.t. the un-normalised embeddings)."""
    _chk(V, labels, dloss, inv_norm)
    B, TF, E = V.shape
    dz = torch.empty_like(V)
    _lib.call("amss_dpcl_loss_bwd_normalized", _p(V), _p(labels), _p(dloss), _p(inv_norm), B, TF, E, S, precision, _p(dz),
              _p(ws), _stream())
    return dz


def dpcl_loss_bwd_normalized_bf16(V, labels, S, dloss, ws, inv_norm):
    """As dpcl_loss_bwd_normalized on the tensor cores, with dz returned as bf16 [B,TF,E] (operand of gemm_bf16)."""
    _chk(V, labels, dloss, inv_norm)
    B, TF, E = V.shape
    dz = torch.empty(B, TF, E, dtype=torch.bfloat16, device=V.device)
    _lib.call("amss_dpcl_loss_bwd_normalized_bf16", _p(V), _p(labels), _p(dloss), _p(inv_norm), B, TF, E, S, _p(dz), _p(ws),
              _stream())
    return dz


def l41_loss_fwd(emb, labels, spk, weights=None):
    """emb[B,TF,E], labels uint8 [B,TF], spk[B,S,E], optional label weights [B,TF] -> loss[1]."""
    _chk(emb, labels, spk, weights)
    B, TF, E = emb.shape
    S = spk.shape[1]
    loss = torch.empty(1, dtype=_f32, device=emb.device)
    ws = _ws(_lib.query("amss_l41_workspace_bytes", B, TF, E, S), emb.device)
    _lib.call("amss_l41_loss_fwd", _p(emb), _p(labels), _p(spk), _p(weights), B, TF, E, S, _p(loss), _p(ws), ws.numel(),
              _stream())
    return loss


def l41_loss_bwd(emb, labels, spk, dloss, weights=None):
    _chk(emb, labels, spk, dloss, weights)
    B, TF, E = emb.shape
    S = spk.shape[1]
    demb = torch.empty_like(emb)
    dspk = torch.empty_like(spk)
    ws = _ws(_lib.query("amss_l41_workspace_bytes", B, TF, E, S), emb.device)
    _lib.call("amss_l41_loss_bwd", _p(emb), _p(labels), _p(spk), _p(weights), _p(dloss), B, TF, E, S, _p(demb), _p(dspk),
              _p(ws), ws.numel(), _stream())
    return demb, dspk


def plugged_labels(front_y, B, S):
    """front_y[B(S+1),Tp,N] -> labels uint8 [B,Tp,N] (argmax_s |X_non_mix|)."""
    _chk(front_y)
    _, Tp, N = front_y.shape
    labels = torch.empty(B, Tp, N, dtype=torch.uint8, device=front_y.device)
    _lib.call("amss_plugged_labels", _p(front_y), B, S, Tp * N, _p(labels), _stream())
    return labels


def wave_stats(target, approx):
    """target, approx [R,L] -> [R,4] = (<t,t>, <a,a>, <t,a>, <t-a,t-a>)."""
    _chk(target, approx)
    R, L = target.shape
    out = torch.empty(R, 4, dtype=_f32, device=target.device)
    _lib.call("amss_wave_stats", _p(target), _p(approx), R, L, _p(out), _stream())
    return out


def wave_stats_bwd(target, approx, dstats):
    """-> d approx [R,L] of wave_stats(target, approx) for the upstream gradient dstats [R,4]."""
    _chk(target, approx, dstats)
    R, L = target.shape
    out = torch.empty_like(approx)
    _lib.call("amss_wave_stats_bwd", _p(target), _p(approx), _p(dstats), R, L, _p(out), _stream())
    return out


def wave_stats_rows(target, approx, approx_div):
    """wave_stats with approx row r // approx_div (target [R,L], approx [R // approx_div, L])."""
    _chk(target, approx)
    R, L = target.shape
    out = torch.empty(R, 4, dtype=_f32, device=target.device)
    _lib.call("amss_wave_stats_rows", _p(target), _p(approx), R, L, int(approx_div), _p(out), _stream())
    return out


def sumsq(tensors):
    """-> device scalar sum of squares of the given tensors (fixed-order reduction, amss_sumsq)."""
    tensors = [t for t in tensors]
    _chk(*tensors)
    out = torch.zeros(1, dtype=_f32, device=tensors[0].device)
    ws = _ws(_lib.query("amss_sumsq_workspace_bytes"), tensors[0].device)
    for t in tensors:
        _lib.call("amss_sumsq", _p(t), t.numel(), _p(out), _p(ws), _stream())
    return out


def adapt_cost_fwd(stats, mix_stats, terms, regsq, B, S, loss_kind, beta, lam, overlap_coef, nonneg_coef):
    """-> (out4 = cost, l2, sdr, sdr_improvement; dstats [B*S,4]; dterms [3]; dreg [1])."""
    _chk(stats, mix_stats, terms, regsq)
    dev = stats.device
    out4 = torch.empty(4, dtype=_f32, device=dev)
    dstats = torch.empty(B * S, 4, dtype=_f32, device=dev)
    dterms = torch.empty(3, dtype=_f32, device=dev)
    dreg = torch.empty(1, dtype=_f32, device=dev)
    _lib.call("amss_adapt_cost_fwd", _p(stats), _p(mix_stats), _p(terms), _p(regsq), B, S, int(loss_kind), float(beta), float(lam),
              float(overlap_coef), float(nonneg_coef), _p(out4), _p(dstats), _p(dterms), _p(dreg), _stream())
    return out4, dstats, dterms, dreg


# ------------------------------------------------------------------------------------------
# k-means
# ------------------------------------------------------------------------------------------
def kmeans_fit(X, init_idx, K, tries, iters, beta=None, notsilent=None, normalize_input=True, assign_at_end=True):
    """X[B,L,E]; init_idx int32 [B*tries,K]; notsilent uint8 [B,L] or None.
    -> (centroids[B,K,E], labels int32 [B,L] or soft [B,L,K], inertia[B,tries], best_try int32 [B])."""
    _chk(X, init_idx, notsilent)
    B, L, E = X.shape
    dev = X.device
    cent = torch.empty(B, K, E, dtype=_f32, device=dev)
    soft = beta is not None
    labels = None if soft else torch.empty(B, L, dtype=torch.int32, device=dev)
    softo = torch.empty(B, L, K, dtype=_f32, device=dev) if soft else None
    inertia = torch.empty(B, tries, dtype=_f32, device=dev)
    best = torch.empty(B, dtype=torch.int32, device=dev)
    ws = _ws(_lib.query("amss_kmeans_workspace_bytes", B, L, E, K, tries), dev)
    _lib.call("amss_kmeans_fit", _p(X), _p(init_idx), _p(notsilent), B, L, E, K, tries, iters,
              float("nan") if beta is None else float(beta), int(normalize_input), int(assign_at_end), _p(cent),
              _p(labels), _p(softo), _p(inertia), _p(best), _p(ws), ws.numel(), _stream())
    return cent, (softo if soft else labels), inertia, best


def kmeans_silence_mask(latent, threshold):
    """latent[B,L] -> notsilent uint8 [B,L] = log10(max/latent) < threshold."""
    _chk(latent)
    B, L = latent.shape
    out = torch.empty(B, L, dtype=torch.uint8, device=latent.device)
    ws = _ws(4 * B, latent.device)
    _lib.call("amss_kmeans_silence_mask", _p(latent), B, L, float(threshold), _p(out), _p(ws), _stream())
    return out


def apply_masks(X_input, S, labels=None, soft=None):
    """X_input[B,TF] -> separated[B*S,TF]."""
    _chk(X_input, labels, soft)
    B, TF = X_input.shape
    out = torch.empty(B * S, TF, dtype=_f32, device=X_input.device)
    _lib.call("amss_apply_masks", _p(X_input), _p(labels), _p(soft), B, S, TF, _p(out), _stream())
    return out


# ------------------------------------------------------------------------------------------
# input contract
# ------------------------------------------------------------------------------------------
def prepare_inputs(x_non_mix, normalize=False):
    """x_non_mix [B,S,L] -> (x_mix [B,L], stats [B*S,2] or None): the mixture as the sequential fp32 sum of the sources;
    normalize=True first normalises every source row IN PLACE to zero mean / unit (population) variance."""
    _chk(x_non_mix)
    B, S, Lw = x_non_mix.shape
    x_mix = torch.empty(B, Lw, dtype=_f32, device=x_non_mix.device)
    stats = torch.empty(B * S, 2, dtype=_f32, device=x_non_mix.device) if normalize else None
    _lib.call("amss_prepare_inputs", _p(x_non_mix), B, S, Lw, int(bool(normalize)), _p(stats), _p(x_mix), _stream())
    return x_mix, stats


def adapt_terms_fwd(y, B, S, rho, separation, want_sep=True):
    """y[B(S+1),Tp,N] -> (sep[B*S,Tp,N] or None, p_hat[Tp*N], terms[3] = sparse_constraint, overlapping, nonneg)."""
    _chk(y)
    TN = y.shape[1] * y.shape[2]
    sep = torch.empty(B * S, y.shape[1], y.shape[2], dtype=_f32, device=y.device) if want_sep else None
    p_hat = torch.empty(TN, dtype=_f32, device=y.device)
    terms = torch.empty(3, dtype=_f32, device=y.device)
    ws = _ws(_lib.query("amss_adapt_terms_workspace_bytes", TN), y.device)
    _lib.call("amss_adapt_terms_fwd", _p(y), B, S, TN, float(rho), int(separation), _p(sep), _p(p_hat), _p(terms), _p(ws),
              ws.numel(), _stream())
    return sep, p_hat, terms


def adapt_terms_bwd(y, p_hat, dsep, dterms, B, S, rho, separation):
    _chk(y, p_hat, dsep, dterms)
    TN = y.shape[1] * y.shape[2]
    dy = torch.empty_like(y)
    _lib.call("amss_adapt_terms_bwd", _p(y), _p(p_hat), _p(dsep), _p(dterms), B, S, TN, float(rho), int(separation), _p(dy),
              _stream())
    return dy


NONLINEARITY = {"softmax": 0, "tanh": 1, "None": 2, None: 2}


def enhance_cost_table(logits, X_input, X_non_mix, nonlinearity="softmax", want_masks=False):
    """logits[B,S,TF], X_input[B,TF], X_non_mix[B,TF,S] -> (table[B,S,S], masks[B,TF,S] or None)."""
    _chk(logits, X_input, X_non_mix)
    B, S, TF = logits.shape
    table = torch.empty(B, S, S, dtype=_f32, device=logits.device)
    masks = torch.empty(B, TF, S, dtype=_f32, device=logits.device) if want_masks else None
    ws = _ws(_lib.query("amss_enhance_cost_workspace_bytes", B, TF, S), logits.device)
    _lib.call("amss_enhance_cost_table", _p(logits), _p(X_input), _p(X_non_mix), B, S, TF, NONLINEARITY[nonlinearity],
              _p(masks), _p(table), _p(ws), ws.numel(), _stream())
    return table, masks


def enhance_cost_bwd(logits, X_input, X_non_mix, perm, dcost_b, nonlinearity="softmax"):
    _chk(logits, X_input, X_non_mix, perm, dcost_b)
    B, S, TF = logits.shape
    dlogits = torch.empty_like(logits)
    _lib.call("amss_enhance_cost_bwd", _p(logits), _p(X_input), _p(X_non_mix), _p(perm), _p(dcost_b), B, S, TF,
              NONLINEARITY[nonlinearity], _p(dlogits), _stream())
    return dlogits


PRE_FUNC = {"None": 0, None: 0, "sqrt": 1, "log": 2}
NORMALIZE = {"None": 0, None: 0, "01": 1, "meanstd": 2}
FUNCTION_MASK = {"None": 0, None: 0, "linear": 1, "sqrt": 2, "square": 3}


def separator_input_prep(X, abs_input=False, pre_func="None", normalize="None", silence_db=0.0):
    """X [B,...] -> same shape: abs -> sqrt / log10 -> '01' / 'meanstd' normalisation over each mixture -> silent-dB mask."""
    _chk(X)
    B = X.shape[0]
    out = torch.empty_like(X)
    _lib.call("amss_separator_input_prep", _p(X), B, X.numel() // B, int(bool(abs_input)), PRE_FUNC[pre_func],
              NORMALIZE[normalize], float(silence_db), _p(out), _stream())
    return out


def label_weights(X, function_mask="None", silence_threshold=0.0):
    """Mixture rows X [B,...] -> label weights [B,TF] (function_mask linear/sqrt/square and/or the silence-loss mask)."""
    _chk(X)
    B = X.shape[0]
    w = torch.empty(B, X.numel() // B, dtype=_f32, device=X.device)
    _lib.call("amss_label_weights", _p(X), B, X.numel() // B, FUNCTION_MASK[function_mask], float(silence_threshold), _p(w),
              _stream())
    return w


# ------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------
def amsgrad_step(p, g, m, v, vhat, lr_t, beta1, beta2, eps, grad_scale=1.0, grad_scale_dev=None):
    _chk(p, g, m, v, vhat, grad_scale_dev)
    _lib.call("amss_amsgrad_step", _p(p), _p(g), _p(m), _p(v), _p(vhat), p.numel(), float(lr_t), float(beta1),
              float(beta2), float(eps), float(grad_scale), _p(grad_scale_dev), _stream())


def momentum_step(p, g, accum, lr, momentum=0.9, grad_scale=1.0, grad_scale_dev=None):
    _chk(p, g, accum, grad_scale_dev)
    _lib.call("amss_momentum_step", _p(p), _p(g), _p(accum), p.numel(), float(lr), float(momentum), float(grad_scale),
              _p(grad_scale_dev), _stream())


def rmsprop_step(p, g, ms, mom, lr, decay=0.9, momentum=0.0, eps=1e-10, grad_scale=1.0, grad_scale_dev=None):
    _chk(p, g, ms, mom, grad_scale_dev)
    _lib.call("amss_rmsprop_step", _p(p), _p(g), _p(ms), _p(mom), p.numel(), float(lr), float(decay), float(momentum),
              float(eps), float(grad_scale), _p(grad_scale_dev), _stream())


def global_norm_clip_factor(gs, clip, grad_scale=1.0):
    """-> device scalar clip / max(||grad_scale * g||, clip)  (tf.clip_by_global_norm); gs: a tensor or a list of
    tensors (the trainable segments of the flat gradient buffer)."""
    gs = [gs] if torch.is_tensor(gs) else list(gs)
    _chk(*gs)
    dev = gs[0].device
    sumsq = torch.zeros(1, dtype=_f32, device=dev)
    ws = _ws(_lib.query("amss_sumsq_workspace_bytes"), dev)
    for g in gs:
        _lib.call("amss_sumsq", _p(g), g.numel(), _p(sumsq), _p(ws), _stream())
    factor = torch.empty(1, dtype=_f32, device=dev)
    _lib.call("amss_clip_factor", _p(sumsq), float(clip), float(grad_scale), _p(factor), _stream())
    return factor


def amsgrad_lr_t(lr, beta1, beta2, step):
    """utils/ops.py:681-683: lr * sqrt(1 - beta2^t) / (1 - beta1^t), t = step (1-based)."""
    return lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
