"""Restatement of the TensorFlow 1.x ops the reference hot path lowers to.

Test infrastructure (see oracle/__init__.py).  Every function takes and returns
torch CPU tensors, is differentiable through torch autograd where the TF op is,
and executes the op sequence TF would (dense conv, dense transposed conv, an
unfused per-step LSTM loop) rather than the optimised GPU algorithm.
All file:line citations are relative to /root/reference.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# small helpers: utils/ops.py:46-59
# --------------------------------------------------------------------------- #
def log10(x):
    """utils/ops.py:56-59 -- log(x) / log(10) (not a native log10)."""
    return torch.log(x) / math.log(10.0)


def logfunc(x, x2):
    """utils/ops.py:46-49 -- x * log(clip(x) / clip(x2)), clip to [1e-10, 1]."""
    x = torch.as_tensor(x, dtype=x2.dtype)
    cx = torch.clamp(x, 1e-10, 1.0)
    cx2 = torch.clamp(x2, 1e-10, 1.0)
    return x * torch.log(cx / cx2)


def kl_div(p, p_hat):
    """utils/ops.py:51-54."""
    return logfunc(p, p_hat) + logfunc(1 - p, 1 - p_hat)


def l2_normalize(x, axis=-1, eps=1e-12):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), eps)).  utils/ops.py:323-324,
    models/Kmeans_2.py:41, models/L41.py:61."""
    ss = (x * x).sum(axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=eps))


def one_hot_argmax(x, depth, on, off, axis=-1):
    """tf.one_hot(tf.argmax(x, axis), depth, on, off): first index wins ties.
    models/network.py:377-378, 501-502."""
    assert axis in (-1, x.dim() - 1)
    idx = torch.argmax(x, dim=-1)  # torch returns the first maximal index
    y = torch.full(x.shape[:-1] + (depth,), float(off), dtype=x.dtype)
    y.scatter_(-1, idx.unsqueeze(-1), float(on))
    return y, idx


# --------------------------------------------------------------------------- #
# adaptive front end: models/adapt.py:95-134
# --------------------------------------------------------------------------- #
def same_pad_1d(L, W, stride):
    """TF 'SAME' padding along one axis -> (out_len, pad_left, pad_right)."""
    out = -(-L // stride)
    pad = max((out - 1) * stride + W - L, 0)
    return out, pad // 2, pad - pad // 2


def conv2d_same_1d(x, filt, stride=1):
    """tf.nn.conv2d(input [Bt,1,L,1], filter [1,W,1,N], strides [1,1,s,1], 'SAME')
    (models/adapt.py:115,119,122).  Cross-correlation:
    X[b,t,n] = sum_k x[b, t*s + k - pad_left] * filt[k,n].
    x: [Bt, L], filt: [W, N] -> [Bt, T, N]."""
    Bt, L = x.shape
    W, N = filt.shape
    _, pl, pr = same_pad_1d(L, W, stride)
    xp = F.pad(x.unsqueeze(1), (pl, pr))
    out = F.conv1d(xp, filt.t().unsqueeze(1), stride=stride)  # [Bt, N, T]
    return out.transpose(1, 2).contiguous()


def max_pool_with_argmax_1d(X, ksize, stride):
    """tf.nn.max_pool_with_argmax(X [Bt,1,L,N], [1,1,ksize,1], [1,1,stride,1], 'VALID')
    (models/adapt.py:116-117).  Returns y [Bt,Tp,N] and the per-sample flat index
    (h*width + w)*C + c = t*N + n as int64 (batch NOT included: that is the only
    convention consistent with utils/ops.py:111-116 re-adding the batch index).
    First maximum wins ties."""
    Bt, L, N = X.shape
    win = X.unfold(1, ksize, stride)          # [Bt, Tp, N, ksize]
    y, rel = win.max(dim=-1)                  # first max index
    Tp = y.shape[1]
    t0 = (torch.arange(Tp) * stride).view(1, Tp, 1)
    n = torch.arange(N).view(1, 1, N)
    argmax = (t0 + rel) * N + n
    return y, argmax.to(torch.int64)


def avg_pool_1d(X, ksize):
    """tf.layers.average_pooling2d(X, (1,ksize), strides=(1,ksize)) 'valid'
    (models/adapt.py:120)."""
    Bt, L, N = X.shape
    Tp = L // ksize
    return X[:, :Tp * ksize].reshape(Bt, Tp, ksize, N).mean(2)


def unpool(pool, ind, L, N):
    """utils/ops.py:94-120 -- scatter_nd of pooled values into [BS, L*N] at the
    per-sample flat index (duplicates add).  pool, ind: [BS, Tp, N] -> [BS, L, N]."""
    BS = pool.shape[0]
    flat = torch.zeros(BS, L * N, dtype=pool.dtype)
    flat = flat.scatter_add(1, ind.reshape(BS, -1), pool.reshape(BS, -1))
    return flat.view(BS, L, N)


def conv2d_transpose_same_1d(U, filt, L, stride=1):
    """tf.nn.conv2d_transpose(U [BS,1,T,N], filter [1,W,1,N], output [BS,1,L,1],
    strides [1,1,s,1], 'SAME') (models/adapt.py:241-243): the exact adjoint of
    conv2d_same_1d:  out[u] = sum_n sum_k U[(u - k + pad_left)/s, n] * filt[k, n].
    U: [BS, T, N] -> [BS, L]."""
    W, N = filt.shape
    _, pl, _ = same_pad_1d(L, W, stride)
    full = F.conv_transpose1d(U.transpose(1, 2), filt.t().unsqueeze(1), stride=stride)
    full = full[:, 0]                                     # [BS, (T-1)*s + W]
    need = pl + L
    if full.shape[1] < need:
        full = F.pad(full, (0, need - full.shape[1]))
    return full[:, pl:pl + L]


# --------------------------------------------------------------------------- #
# STFT twin: models/network.py:480-502, 584-607
# --------------------------------------------------------------------------- #
def hann_periodic(n, dtype=torch.float32):
    """tf.contrib.signal.hann_window(n, periodic=True)."""
    k = torch.arange(n, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).to(dtype)


def stft(x, frame_length, frame_step):
    """tf.contrib.signal.stft(x, frame_length, frame_step, fft_length=frame_length)
    (models/network.py:482-492): pad_end=False, periodic hann, rFFT.
    x: [R, L] -> complex [R, T, frame_length//2+1], T = 1 + (L - frame_length)//frame_step."""
    frames = x.unfold(-1, frame_length, frame_step)
    frames = frames * hann_periodic(frame_length, x.dtype)
    return torch.fft.rfft(frames, n=frame_length)


def inverse_stft_window(frame_length, frame_step, dtype=torch.float32):
    """tf.contrib.signal.inverse_stft_window_fn(frame_step)(frame_length)
    (models/network.py:602): w / tile(sum over overlaps of w^2)."""
    w = hann_periodic(frame_length, torch.float64)
    denom = w * w
    overlaps = -(-frame_length // frame_step)
    denom = F.pad(denom, (0, overlaps * frame_step - frame_length))
    denom = denom.view(overlaps, frame_step).sum(0, keepdim=True).repeat(overlaps, 1).reshape(-1)
    return (w / denom[:frame_length]).to(dtype)


def overlap_and_add(frames, frame_step):
    """tf.contrib.signal.overlap_and_add. frames [R, T, n] -> [R, (T-1)*step + n]."""
    R, T, n = frames.shape
    out_len = (T - 1) * frame_step + n
    out = torch.zeros(R, out_len, dtype=frames.dtype)
    for t in range(T):
        out[:, t * frame_step:t * frame_step + n] += frames[:, t]
    return out


def inverse_stft(stfts, frame_length, frame_step):
    """tf.contrib.signal.inverse_stft(stfts, frame_length, frame_step,
    window_fn=inverse_stft_window_fn(frame_step)) (models/network.py:598-602)."""
    real = torch.fft.irfft(stfts, n=frame_length)[..., :frame_length]
    real = real * inverse_stft_window(frame_length, frame_step, real.dtype)
    return overlap_and_add(real, frame_step)


# --------------------------------------------------------------------------- #
# BLSTM: utils/ops.py:358-383
# --------------------------------------------------------------------------- #
def basic_lstm_rnn(x, kernel, bias, forget_bias=1.0):
    """tf.nn.dynamic_rnn(BasicLSTMCell(H), x) with zero initial state.
    x [B,T,I]; kernel [I+H, 4H]; bias [4H].  Gate split order i, j, f, o;
    c' = c*sigmoid(f + forget_bias) + sigmoid(i)*tanh(j); h' = tanh(c')*sigmoid(o).
    DropoutWrapper with keep_prob 1.0 (drop_val 0.0, utils/trainer.py:77-78) is
    the identity and is not restated."""
    B, T, I = x.shape
    H = kernel.shape[1] // 4
    h = torch.zeros(B, H, dtype=x.dtype)
    c = torch.zeros(B, H, dtype=x.dtype)
    outs = []
    for t in range(T):
        z = torch.cat([x[:, t], h], 1) @ kernel + bias
        i, j, f, o = z.split(H, dim=1)
        c = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
        h = torch.tanh(c) * torch.sigmoid(o)
        outs.append(h)
    return torch.stack(outs, 1)


def blstm(x, kernel_fw, bias_fw, kernel_bw, bias_bw):
    """BLSTM.f_prop (utils/ops.py:366-383): forward LSTM on x, backward LSTM on
    reverse(x, time), concat([fwd, reverse(bwd)], 2)."""
    fw = basic_lstm_rnn(x, kernel_fw, bias_fw)
    bw = basic_lstm_rnn(torch.flip(x, [1]), kernel_bw, bias_bw)
    return torch.cat([fw, torch.flip(bw, [1])], 2)


def conv1d_k1(x, W, b):
    """Conv1D with filter [1, in, out] (utils/ops.py:486-503): per-frame dense."""
    return x @ W + b
