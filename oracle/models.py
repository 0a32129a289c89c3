"""Restatement of the reference model graphs (Adapt, Separator, DPCL, L41).

Test infrastructure (see oracle/__init__.py).  Parameters are plain dicts of
torch CPU tensors keyed by the reference's TF variable names:
  front/window/w [W], front/bases/bases [W,N], back/window/value, back/bases/value,
  prediction/{forward,backward}_BLSTM_i/rnn/basic_lstm_cell/{kernel,bias},
  prediction/W [1,C,E*F] (stored [C,E*F]), prediction/b, speaker_centroids [251,E],
  enhance/{forward,backward}_BLSTM_i/..., enhance/W, enhance/b.
All file:line citations are relative to /root/reference.
"""
import itertools
import math

import numpy as np
import torch

from . import tf_ops as T


# --------------------------------------------------------------------------- #
# parameter construction (distribution-faithful; TF's RNG stream cannot be matched)
# --------------------------------------------------------------------------- #
def _glorot(gen, shape, fan_in, fan_out, dtype):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(dtype)


def init_adapt_params(window, filters, seed=42, dtype=torch.float32):
    """models/adapt.py:104-105, 232-233: xavier_initializer_conv2d variables."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for scope, wn, bn in (("front", "window/w", "bases/bases"), ("back", "window/value", "bases/value")):
        p[f"{scope}/{wn}"] = _glorot(g, (window,), window, 1, dtype)
        p[f"{scope}/{bn}"] = _glorot(g, (window, filters), window, filters, dtype)
    return p


def init_blstm_stack(prefix, nb_layers, in_dim, layer_size, gen, dtype):
    """utils/ops.py:366-383: BasicLSTMCell(layer_size//2) per direction, kernel
    glorot-uniform [I+H,4H], bias zeros.  Layer 0 name has no suffix in TF's
    uniquified scopes only for the enclosing @scope; the per-layer names are
    forward_BLSTM_i / backward_BLSTM_i."""
    p = {}
    H = layer_size // 2
    I = in_dim
    for i in range(nb_layers):
        for d in ("forward", "backward"):
            base = f"{prefix}/{d}_BLSTM_{i}/rnn/basic_lstm_cell"
            p[base + "/kernel"] = _glorot(gen, (I + H, 4 * H), I + H, 4 * H, dtype)
            p[base + "/bias"] = torch.zeros(4 * H, dtype=dtype)
        I = 2 * H
    return p


def init_conv1d(prefix, in_dim, out_dim, gen, dtype, reference_scale=False):
    """utils/ops.py:487-495.  The reference draws U(+-sqrt(2/sqrt(2/(in+out)))) (~+-12
    for 600->10280, a quirk of :489-494); reference_scale=True reproduces that
    range, the default uses a glorot range so synthetic runs stay well conditioned."""
    if reference_scale:
        fan = math.sqrt(2.0 / float(in_dim + out_dim))
        lim = math.sqrt(2.0 / fan)
    else:
        lim = math.sqrt(6.0 / (in_dim + out_dim))
    W = ((torch.rand((in_dim, out_dim), generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(dtype)
    return {prefix + "/W": W, prefix + "/b": torch.zeros(out_dim, dtype=dtype)}


def init_separator_params(F_bins, nb_layers, layer_size, embedding_size, seed=42, dtype=torch.float32,
                          tot_speakers=251, with_speaker_vectors=False, reference_scale=False):
    g = torch.Generator().manual_seed(seed + 1)
    p = init_blstm_stack("prediction", nb_layers, F_bins, layer_size, g, dtype)
    p.update(init_conv1d("prediction", 2 * (layer_size // 2), embedding_size * F_bins, g, dtype, reference_scale))
    if with_speaker_vectors:
        # models/L41.py:16-18 truncated_normal(stddev=sqrt(2/E))
        std = math.sqrt(2.0 / embedding_size)
        v = torch.randn(tot_speakers, embedding_size, generator=g, dtype=torch.float64)
        v = torch.clamp(v, -2.0, 2.0) * std
        p["speaker_centroids"] = v.to(dtype)
    return p


def init_enhance_params(F_bins, nb_layers, layer_size, seed=42, dtype=torch.float32):
    """models/network.py:629-636."""
    g = torch.Generator().manual_seed(seed + 2)
    p = init_blstm_stack("enhance", nb_layers, 2 * F_bins, layer_size, g, dtype)
    p.update(init_conv1d("enhance", 2 * (layer_size // 2), F_bins, g, dtype))
    return p


# --------------------------------------------------------------------------- #
# Network.sdr_improvement: models/network.py:196-221
# --------------------------------------------------------------------------- #
def sdr_improvement(x_mix, s_target, s_approx):
    """Returns (SDR improvement scalar, 'sdr' loss [B,S]).  with_perm=False branch."""
    S = s_target.shape[1]
    mix = x_mix.unsqueeze(1).repeat(1, S, 1)
    tn = (s_target ** 2).sum(-1)
    an = (s_approx ** 2).sum(-1)
    mn = (mix ** 2).sum(-1)
    ts2 = ((s_target * s_approx).sum(-1)) ** 2
    tm2 = ((s_target * mix).sum(-1)) ** 2
    sep = 1.0 / ((tn * an) / ts2 - 1.0)
    separated = 10.0 * T.log10(sep)
    non_separated = 10.0 * T.log10(1.0 / ((tn * mn) / tm2 - 1.0))
    loss = (tn * an) / (ts2 + 1e-12)
    val = (separated - non_separated).mean(-1).mean(-1)
    return val, loss


# --------------------------------------------------------------------------- #
# Adapt: models/adapt.py
# --------------------------------------------------------------------------- #
def adapt_filters(p, scope):
    """models/adapt.py:106 / :234: |window[:,None]| * bases -> [W, N]."""
    if scope == "front":
        return p["front/window/w"].abs().unsqueeze(1) * p["front/bases/bases"]
    return p["back/window/value"].abs().unsqueeze(1) * p["back/bases/value"]


def adapt_front(p, x_mix, x_non_mix, max_pool, hop, with_max_pool=True, with_average_pool=False,
                sparsity=0.01):
    """models/adapt.py:41-48 + 95-134.  Returns dict with y [Btot,Tp,N], argmax
    (int64, max-pool mode only), p_hat [Tp*N], sparse_constraint (scalar)."""
    B, S, L = x_non_mix.shape
    x = torch.cat([x_mix, x_non_mix.reshape(B * S, L)], 0)
    filt = adapt_filters(p, "front")
    argmax = None
    if with_max_pool:
        X = T.conv2d_same_1d(x, filt, 1)
        y, argmax = T.max_pool_with_argmax_1d(X, max_pool, hop)
    elif with_average_pool:
        X = T.conv2d_same_1d(x, filt, 1)
        y = T.avg_pool_1d(X, max_pool)
    else:
        y = T.conv2d_same_1d(x, filt, hop)
    p_hat = y.abs().reshape(y.shape[0], -1).sum(0)
    sparse = T.kl_div(sparsity, p_hat).sum()
    return {"y": y, "argmax": argmax, "p_hat": p_hat, "sparse_constraint": sparse, "filt": filt}


def adapt_overlap(y, B, S):
    """models/adapt.py:141-160: 1 - |a-b| / (max(a,b) + 1e-8), mean over bins,
    speaker pairs and batch, on |front output| of the non-mixed rows."""
    Tp, N = y.shape[1], y.shape[2]
    nm = y[B:].reshape(B, S, Tp * N).abs()
    vals = []
    for a, b in itertools.combinations(range(S), 2):
        pa, pb = nm[:, a], nm[:, b]
        measure = 1.0 - (pa - pb).abs() / (torch.maximum(pa, pb) + 1e-8)
        vals.append(measure.mean(-1))
    return torch.stack(vals, 1).mean(-1).mean(-1)


def adapt_separator_pretraining(y, B, S, separation="perfect"):
    """models/adapt.py:162-196.  y [B(S+1),Tp,N] -> [B*S,Tp,N]."""
    Tp, N = y.shape[1], y.shape[2]
    input_mix = y[:B].reshape(B, 1, Tp, N).repeat(1, S, 1, 1)
    input_non_mix = y[B:].reshape(B, S, Tp, N)
    if separation == "mask":
        filters = input_non_mix / input_mix
        out = input_mix * filters
    else:  # 'perfect'
        tiled_sum = input_non_mix.sum(1, keepdim=True).repeat(1, S, 1, 1)
        out = input_mix - (tiled_sum - input_non_mix)
    return out.reshape(B * S, Tp, N)


def adapt_back(p, sep_out, argmax, B, S, L, hop, with_max_pool=True, with_average_pool=False, max_pool=None):
    """models/adapt.py:205-252.  sep_out [B*S,Tp,N]; argmax [B(S+1),Tp,N] (only the
    mixture rows [:B] are used, tiled S x: :212-218).  -> [B,S,L]."""
    N = sep_out.shape[2]
    filt2 = adapt_filters(p, "back")
    if with_max_pool:
        am = argmax[:B].unsqueeze(1).repeat(1, S, 1, 1).reshape(B * S, -1, N)
        U = T.unpool(sep_out, am, L, N)
        out = T.conv2d_transpose_same_1d(U, filt2, L, 1)
    elif with_average_pool:
        U = sep_out.repeat_interleave(max_pool, dim=1)  # UpSampling2D((1, pool))
        if U.shape[1] < L:
            U = torch.nn.functional.pad(U, (0, 0, 0, L - U.shape[1]))
        out = T.conv2d_transpose_same_1d(U, filt2, L, 1)
    else:
        out = T.conv2d_transpose_same_1d(sep_out, filt2, L, hop)
    return out.reshape(B, S, L), filt2


def adapt_pretraining_cost(p, x_mix, x_non_mix, *, max_pool, hop, loss="sdr", separation="perfect",
                           beta=1e-2, regularization=1e-4, sparsity=0.01, overlap_coef=1e-3,
                           non_negativity=0.0, with_max_pool=True, with_average_pool=False):
    """Adapt.cost, pretraining branch (models/adapt.py:307-338, 374-385).
    Reproduces the doubled lambda (:312 and :380) and doubled non_negativity
    (:316 and :384) factors.  Returns (cost, aux dict)."""
    B, S, L = x_non_mix.shape
    fr = adapt_front(p, x_mix, x_non_mix, max_pool, hop, with_max_pool, with_average_pool, sparsity)
    y = fr["y"]
    overlapping = adapt_overlap(y, B, S)
    sep = adapt_separator_pretraining(y, B, S, separation)
    back, filt2 = adapt_back(p, sep, fr["argmax"], B, S, L, hop, with_max_pool, with_average_pool, max_pool)

    reg = regularization * (0.5 * (filt2 ** 2).sum() + 0.5 * (fr["filt"] ** 2).sum())
    neg = torch.where(y < 0, y, torch.zeros_like(y)) ** 2
    nn = non_negativity * neg.reshape(neg.shape[0], -1).sum(1).mean()

    l2 = ((x_non_mix - back) ** 2).sum(-1).sum(-1).mean(-1)
    sdr_imp, sdr = sdr_improvement(x_mix, x_non_mix, back)
    sdr = sdr.mean(-1).mean(-1)
    if loss == "l2":
        cost = l2
    elif loss == "sdr":
        cost = sdr
    else:
        cost = l2 + sdr
    if beta != 0.0:
        cost = cost + beta * fr["sparse_constraint"]
    if regularization != 0.0:
        cost = cost + regularization * reg
    if overlap_coef != 0.0:
        cost = cost + overlap_coef * overlapping
    if non_negativity is not None:
        cost = cost + non_negativity * nn
    aux = {"y": y, "argmax": fr["argmax"], "back": back, "l2": l2, "sdr": sdr, "sdr_improvement": sdr_imp,
           "sparse_constraint": fr["sparse_constraint"], "overlapping": overlapping, "p_hat": fr["p_hat"]}
    return cost, aux


# --------------------------------------------------------------------------- #
# Separator: models/network.py:313-723
# --------------------------------------------------------------------------- #
def separator_preprocessing(x_mix, x_non_mix, window_size, hop_size, a, b):
    """Separator.preprocessing (models/network.py:480-502).  Returns stfts (complex
    [B,T,F]), X=|stft| [B,T,F], X_non_mix [B,T,F,S], y [B,T,F,S] (one_hot on/off a/b),
    argmax [B,T,F]."""
    B, S, L = x_non_mix.shape
    stfts = T.stft(x_mix, window_size, hop_size)
    st_nm = T.stft(x_non_mix.reshape(B * S, L), window_size, hop_size)
    X = stfts.abs()
    Fb = X.shape[-1]
    X_nm = st_nm.abs().reshape(B, S, -1, Fb).permute(0, 2, 3, 1)
    y, idx = T.one_hot_argmax(X_nm, S, a, b)
    return {"stfts": stfts, "X": X, "X_non_mix": X_nm, "y": y, "argmax": idx}


def separator_plugged_inputs(front_y, B, S, a, b):
    """Separator.__init__ plugged branch (models/network.py:357-400), default flags
    (function_mask 'None', silence_loss False).  front_y [B(S+1),Tp,N]."""
    X = front_y[:B]
    X_nm = front_y[B:].reshape(B, S, front_y.shape[1], front_y.shape[2]).permute(0, 2, 3, 1)
    y, idx = T.one_hot_argmax(X_nm.abs(), S, a, b)
    return {"X": X, "X_non_mix": X_nm, "y": y, "argmax": idx}


def separator_input_prep(X, plugged, abs_input=False, pre_func="None", normalize="None", silence_db=0.0):
    """Separator.init_separator (models/network.py:409-443) + normalization01 / normalization_mean_std (:504-521).
    plugged: abs_input -> normalisation.  STFT: pre_func (sqrt | log10(x + 1e-12)) -> normalisation -> silent-dB mask
    `mask = (max - X) < silence_db / 20` (max over (T,F) of the normalised X)."""
    if plugged:
        if abs_input:
            X = X.abs()
    else:
        if pre_func == "sqrt":
            X = torch.sqrt(X)
        elif pre_func == "log":
            X = T.log10(X + 1e-12)
    if normalize == "01":
        mn = X.amin((1, 2), keepdim=True)
        mx = X.amax((1, 2), keepdim=True)
        X = (X - mn) / (mx - mn)
    elif normalize == "meanstd":
        mean = X.mean((1, 2), keepdim=True)
        var = X.var((1, 2), unbiased=False, keepdim=True)
        X = (X - mean) / torch.sqrt(var)
    if not plugged and silence_db > 0:
        mx = X.amax((1, 2), keepdim=True)
        X = ((mx - X) < silence_db / 20.0).to(X.dtype) * X
    return X


def plugged_label_weights(X, function_mask="None", silence_loss=False, threshold_silence_loss=2.0):
    """Separator.__init__ plugged branch (models/network.py:381-396): the factor the one-hot labels are multiplied by:
    function_mask linear |X|/max, sqrt, square; silence_loss: [log10(max/|X|) < threshold].  X = mixture rows [B,T,N]."""
    a = X.abs()
    mx = a.amax((1, 2), keepdim=True)
    w = torch.ones_like(a)
    if function_mask == "linear":
        w = a / mx
    elif function_mask == "sqrt":
        w = torch.sqrt(a / mx)
    elif function_mask == "square":
        w = (a / mx) ** 2
    if silence_loss:
        w = w * (T.log10(mx / a) < threshold_silence_loss).to(a.dtype)
    return w


def blstm_stack(p, prefix, nb_layers, x):
    for i in range(nb_layers):
        f = f"{prefix}/forward_BLSTM_{i}/rnn/basic_lstm_cell"
        bk = f"{prefix}/backward_BLSTM_{i}/rnn/basic_lstm_cell"
        x = T.blstm(x, p[f + "/kernel"], p[f + "/bias"], p[bk + "/kernel"], p[bk + "/bias"])
    return x


def separator_prediction(p, X, nb_layers, embedding_size, normalize=True):
    """DPCL.prediction (models/dpcl.py:19-39) / L41Model.prediction (models/L41.py:21-45):
    N x BLSTM -> Conv1D(1x1) -> Reshape [B,T,F,E] -> (L2 normalise axis 3)."""
    B, Tt, Fb = X.shape
    h = blstm_stack(p, "prediction", nb_layers, X)
    z = T.conv1d_k1(h, p["prediction/W"], p["prediction/b"]).reshape(B, Tt, Fb, embedding_size)
    return T.l2_normalize(z, 3) if normalize else z


def dpcl_cost(V4, y):
    """DPCL.cost (models/dpcl.py:41-86): un-squared Frobenius norms, mean over batch."""
    B, Tt, Fb, S = y.shape
    E = V4.shape[-1]
    Y = y.reshape(B, Tt * Fb, S)
    V = V4.reshape(B, Tt * Fb, E)
    ones = torch.ones(B, Tt * Fb, 1, dtype=V.dtype)
    mul_ones = Y.transpose(1, 2) @ ones
    diagonal = Y @ mul_ones
    D = (1.0 / torch.sqrt(diagonal)).reshape(B, Tt * Fb)
    DV = D.unsqueeze(-1) * V
    VTV = V.transpose(1, 2) @ DV
    DY = D.unsqueeze(-1) * Y
    VTY = V.transpose(1, 2) @ DY
    YTY = Y.transpose(1, 2) @ DY
    fro = lambda M: torch.sqrt((M * M).sum((-2, -1)))
    cost = fro(VTV) - 2 * fro(VTY) + fro(YTY)
    return cost.mean()


def l41_cost(p, emb, y, I, normalize=True):
    """L41Model.cost, sampling=None branch (models/L41.py:47-63, 150-178).  y may already carry the label weights of
    --function_mask / --silence_loss (y * w[..., None], models/network.py:381-396)."""
    sv = p["speaker_centroids"]
    if normalize:
        sv = T.l2_normalize(sv, 1)
    Vspk = sv[I.long()]                                   # [B,S,E]
    dot = (Vspk[:, None, None, :, :] * emb[:, :, :, None, :]).sum(4)   # [B,T,F,S]
    cost = -torch.log(torch.sigmoid(y * dot))
    cost = cost.mean(3).mean(0).mean()
    return cost


def separate(V4, X_input, kmeans_fn, S, beta=None):
    """Separator.separate (models/network.py:554-582).  kmeans_fn(embeddings [B,TF,E])
    -> labels ([B,TF] int for hard, [B,TF,S] float for soft).  Returns
    (separated [B*S,T,F], masks [B,TF,S])."""
    B, Tt, Fb, E = V4.shape
    labels = kmeans_fn(V4.reshape(B, Tt * Fb, E))
    if beta is None:
        masks = torch.nn.functional.one_hot(labels.long(), S).to(X_input.dtype)
    else:
        masks = labels
    sep = X_input.reshape(B, -1, 1) * masks
    sep = sep.reshape(B, Tt, Fb, S).permute(0, 3, 1, 2).reshape(B * S, Tt, Fb)
    return sep, masks


def postprocessing(separated, stfts, S, window_size, hop_size):
    """Separator.postprocessing (models/network.py:584-607): mask*|X| * exp(j*angle(X_mix))
    -> inverse_stft -> [B,S,L']."""
    B = stfts.shape[0]
    angles = torch.angle(stfts).unsqueeze(1).repeat(1, S, 1, 1).reshape(separated.shape)
    spec = torch.complex(separated, torch.zeros_like(separated)) * torch.exp(torch.complex(torch.zeros_like(angles), angles))
    out = T.inverse_stft(spec, window_size, hop_size)
    return out.reshape(B, S, -1)


def enhance(p, separated, X_input, S, nb_layers_enhance, normalize_enhance=False, nonlinearity="softmax"):
    """Separator.enhance (models/network.py:610-660).  separated [B*S,T,F], X_input [B,T,F].
    Returns (enhanced [B,S,TF] == self.separated, cost_in [B,TF,S], masks [B,TF,S])."""
    B, Tt, Fb = X_input.shape
    sep4 = separated.reshape(B, S, Tt, Fb)
    X_in = X_input.unsqueeze(1).repeat(1, S, 1, 1)
    z = torch.cat([sep4, X_in], 3).reshape(B * S, Tt, 2 * Fb)
    if normalize_enhance:
        mean = z.mean((1, 2), keepdim=True)
        var = z.var((1, 2), unbiased=False, keepdim=True)
        z = (z - mean) / torch.sqrt(var)
    h = blstm_stack(p, "enhance", nb_layers_enhance, z)
    yv = T.conv1d_k1(h, p["enhance/W"], p["enhance/b"])          # [B*S,T,F]
    yv = yv.reshape(B, S, Tt * Fb).transpose(1, 2)               # [B,TF,S]
    if nonlinearity == "softmax":
        yv = torch.softmax(yv, -1)
    elif nonlinearity == "tanh":
        yv = torch.tanh(yv)
    masks = yv
    cost_in = yv * X_input.reshape(B, -1, 1)
    return cost_in.transpose(1, 2), cost_in, masks


def enhance_cost(cost_in, X_non_mix):
    """Separator.enhance_cost (models/network.py:662-693): PIT L2 over S! perms."""
    B, TF, S = cost_in.shape
    est = cost_in.transpose(1, 2)                                 # [B,S,TF]
    tgt = X_non_mix.reshape(B, TF, S).transpose(1, 2)             # [B,S,TF]
    costs = []
    for perm in itertools.permutations(range(S)):
        costs.append(((tgt - est[:, list(perm)]) ** 2).sum(-1).sum(-1))
    return torch.stack(costs, 1).min(1).values.mean()


def adapt_separation_cost(p, x_mix, x_non_mix, back, *, loss="sdr", regularization=1e-4):
    """Adapt.cost, pretraining=False branch (models/adapt.py:339-372) + Network.sdr_improvement(with_perm=True)
    (models/network.py:196-221), written with the SAME broadcasting the TF graph performs: the targets are reshaped to
    [B,1,S,L] while `back` stays [B,S,L] (= [1,B,S,L]), so every product below has shape [B,B,S(,L)] -- targets of mixture
    b against the estimates of every mixture b' -- and `reduce_min(sdr, 1)` runs over b'.  Only the l2 term sees the
    permuted estimates.  KL / overlap / non-negativity terms need the front output and are left to the caller."""
    B, S, L = x_non_mix.shape
    perms = list(itertools.permutations(range(S)))
    permuted = torch.stack([back[:, list(pm)] for pm in perms], 1)          # [B,P,S,L]
    X_nmr = x_non_mix.reshape(B, 1, S, L)
    l2 = ((X_nmr - permuted) ** 2).mean(-1).sum(-1).min(-1).values.mean(-1)
    # sdr_improvement(X_nmr, back, True)
    mix = x_mix.unsqueeze(1).repeat(1, S, 1)                                # [B,S,L]
    tn = (X_nmr ** 2).sum(-1)                                               # [B,1,S]
    an = (back ** 2).sum(-1)                                                # [B,S]   -> broadcasts as [1,B,S]
    mn = (mix ** 2).sum(-1)
    ts2 = ((X_nmr * back).sum(-1)) ** 2                                     # [B,B,S]
    tm2 = ((X_nmr * mix).sum(-1)) ** 2
    separated = 10.0 * T.log10(1.0 / ((tn * an) / ts2 - 1.0))
    non_separated = 10.0 * T.log10(1.0 / ((tn * mn) / tm2 - 1.0))
    sdr_tab = (tn * an) / (ts2 + 1e-12)
    val = (separated - non_separated).mean(-1).mean(0).max(-1).values
    sdr = sdr_tab.min(1).values.sum(-1).mean(-1)
    cost = l2 if loss == "l2" else (sdr if loss == "sdr" else 1e-3 * l2 + sdr)
    if regularization != 0.0:
        f1, f2 = adapt_filters(p, "front"), adapt_filters(p, "back")
        cost = cost + regularization * (regularization * (0.5 * (f2 ** 2).sum() + 0.5 * (f1 ** 2).sum()))
    return cost, {"l2": l2, "sdr": sdr, "sdr_improvement": val}


def cost_finetuning(x_non_mix, est):
    """cost_finetuning (models/network.py:697-723, models/adapt.py:404-431):
    0.5*sum_L (x - xhat)^2, mean over S, min over perms, mean over B."""
    B, S, L = x_non_mix.shape
    costs = []
    for perm in itertools.permutations(range(S)):
        costs.append((0.5 * ((x_non_mix - est[:, list(perm)]) ** 2).sum(-1)).mean(-1))
    return torch.stack(costs, 1).min(1).values.mean()


# --------------------------------------------------------------------------- #
# synthetic "LibriSpeech-shaped" mixtures: SURVEY.md section 8(d)
# --------------------------------------------------------------------------- #
def synthetic_mixtures(B, S, L, seed=42, fs=16000, tot_speakers=251):
    """Seeded speech-like sources: 12 harmonics of a slowly varying f0 in [90,250] Hz with
    1/k roll-off + low-passed noise, a 3-6 Hz syllabic envelope with ~25 % silence, RMS 0.05.
    mix = sum of sources (data/dataset.py:462-468); distinct speaker ids per mixture
    (data/dataset.py:473-480).  Returns float32 numpy (x_mix [B,L], x_non_mix [B,S,L], I [B,S])."""
    rng = np.random.RandomState(seed)
    t = np.arange(L) / float(fs)
    src = np.zeros((B, S, L), np.float64)
    for b in range(B):
        for s in range(S):
            f0 = rng.uniform(90, 250)
            vib = 1.0 + 0.05 * np.sin(2 * np.pi * rng.uniform(0.5, 2.0) * t + rng.uniform(0, 6.28))
            phase = 2 * np.pi * np.cumsum(f0 * vib) / fs
            sig = np.zeros(L)
            for k in range(1, 13):
                sig += np.sin(k * phase + rng.uniform(0, 6.28)) / k
            noise = rng.randn(L)
            noise = np.convolve(noise, np.ones(8) / 8.0, mode="same")
            sig = sig + 0.3 * noise
            env = 0.5 * (1 + np.sin(2 * np.pi * rng.uniform(3, 6) * t + rng.uniform(0, 6.28)))
            env = np.clip((env - 0.25) / 0.75, 0.0, 1.0)
            sig = sig * env
            sig *= 0.05 / (np.sqrt(np.mean(sig ** 2)) + 1e-12)
            src[b, s] = sig
    I = np.stack([rng.choice(tot_speakers, size=S, replace=False) for _ in range(B)]).astype(np.int32)
    x_non_mix = src.astype(np.float32)
    return x_non_mix.sum(1).astype(np.float32), x_non_mix, I
