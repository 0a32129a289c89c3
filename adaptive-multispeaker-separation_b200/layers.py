"""The reference's layer protocol on top of the CUDA library: any object with ``f_prop(x)``,
composed by ``f_props(layers, x)`` (reference utils/ops.py:82-85), plus the parameter store that
keeps every trainable tensor inside ONE flat fp32 buffer (so AMSGrad is a single fused kernel and
the data-parallel gradient exchange is a single NCCL all-reduce).

Each layer's forward/backward is a kernel of libamss_b200.so wrapped in a torch.autograd.Function;
torch only records the tape.  Variable names follow the reference's TF checkpoint names
(SURVEY.md section 5): ``prediction/forward_BLSTM_0/rnn/basic_lstm_cell/kernel`` etc.
"""
import math

import torch

from . import ops
from ._lib import AMSS_PREC_FP32


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
class ParamStore:
    """Flat parameter / gradient buffers with named views.

    register() before finalize(); after finalize() ``self[name]`` is a torch.nn.Parameter that
    views ``self.flat`` and whose ``.grad`` views ``self.grad_flat``."""

    def __init__(self, device="cuda", seed=42):
        self.device = torch.device(device)
        self.gen = torch.Generator().manual_seed(seed)      # reference: seed 42 (config.py:7)
        self._specs = []            # (name, shape, init tensor (cpu), trainable)
        self.params = {}
        self.flat = None
        self.grad_flat = None
        self._slices = {}
        self._segments = None

    def register(self, name, init, trainable=True):
        if self.flat is not None:
            raise RuntimeError("ParamStore already finalized")
        if any(n == name for n, *_ in self._specs):
            raise KeyError(f"duplicate parameter {name}")
        self._specs.append((name, tuple(init.shape), init.to(torch.float32), trainable))

    def finalize(self):
        # trainable parameters first so that the optimizer / all-reduce see one contiguous range
        specs = [s for s in self._specs if s[3]] + [s for s in self._specs if not s[3]]
        off = 0
        self.n_trainable = 0
        for name, shape, init, tr in specs:
            n = int(math.prod(shape))
            self._slices[name] = (off, n, shape, tr)
            off += (n + 3) // 4 * 4                          # keep every view 16-byte aligned
            if tr:
                self.n_trainable = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grad_flat = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
        for name, shape, init, tr in specs:
            o, n, _, _ = self._slices[name]
            self.flat[o:o + n].copy_(init.reshape(-1))
            p = torch.nn.Parameter(self.flat[o:o + n].view(shape), requires_grad=tr)
            if tr:
                p.grad = self.grad_flat[o:o + n].view(shape)
            self.params[name] = p
        return self

    def __getitem__(self, name):
        return self.params[name]

    def names(self):
        return list(self.params)

    def set_trainable(self, predicate):
        """Freeze / unfreeze by name (the reference's freeze_all_with / --train substrings).
        Frozen parameters keep their slot in the flat buffer, but leave the trainable segments: the optimizer and the
        gradient exchange skip them (the reference removes them from the optimizer's var_list), and their gradient
        slice is zeroed so that nothing stale is ever applied if they are unfrozen again."""
        for name, p in self.params.items():
            o, n, shape, tr = self._slices[name]
            if not tr:
                continue
            keep = bool(predicate(name))
            if p.requires_grad and not keep:
                self.grad_flat[o:o + n].zero_()
            p.requires_grad_(keep)
        self._segments = None

    def trainable_segments(self):
        """[(offset, length)] of the maximal contiguous runs of the flat buffer occupied by parameters that currently
        require a gradient (16-byte aligned by construction; the alignment padding between neighbours is included)."""
        if self._segments is None:
            segs = []
            for name, (o, n, _, tr) in sorted(self._slices.items(), key=lambda kv: kv[1][0]):
                if not tr or not self.params[name].requires_grad:
                    continue
                n4 = (n + 3) // 4 * 4
                if segs and segs[-1][0] + segs[-1][1] == o:
                    segs[-1] = (segs[-1][0], segs[-1][1] + n4)
                else:
                    segs.append((o, n4))
            self._segments = segs
        return self._segments

    def grad_buckets(self):
        """[(lo, hi, [names])] : the trainable parameters grouped per layer (name up to the last '/'-component that
        identifies the layer), contiguous in the flat buffer, for the bucketed gradient exchange (dp.GradBuckets)."""
        out = []
        for name, (o, n, _, tr) in sorted(self._slices.items(), key=lambda kv: kv[1][0]):
            if not tr or not self.params[name].requires_grad:
                continue
            n4 = (n + 3) // 4 * 4
            parts = name.split("/")
            key = parts[0] + "/" + parts[1].replace("forward_", "").replace("backward_", "") if len(parts) > 2 else parts[0]
            if out and out[-1][3] == key and out[-1][1] == o:
                out[-1] = (out[-1][0], o + n4, out[-1][2] + [name], key)
            else:
                out.append((o, o + n4, [name], key))
        return [(lo, hi, names) for lo, hi, names, _ in out]

    def trainable_span(self):
        """(lo, hi) covering every trainable segment: the range the single gradient all-reduce runs over."""
        segs = self.trainable_segments()
        return (segs[0][0], segs[-1][0] + segs[-1][1]) if segs else (0, 0)

    def state_dict(self):
        return {k: v.detach().cpu().clone() for k, v in self.params.items()}

    def load_state_dict(self, sd, strict=True):
        for k, v in sd.items():
            if k not in self.params:
                if strict:
                    raise KeyError(k)
                continue
            v = torch.as_tensor(v)
            tgt = self.params[k]
            if tuple(v.shape) != tuple(tgt.shape):
                # a raw TF checkpoint stores Conv1D filters as [1, in, out] (utils/ops.py:486-494): singleton axes may differ,
                # anything else (which copy_ would silently broadcast) is an error
                squeeze = lambda sh: tuple(d for d in sh if d != 1)  # noqa: E731
                if squeeze(v.shape) != squeeze(tgt.shape):
                    raise ValueError(f"{k}: stored shape {tuple(v.shape)} does not match {tuple(tgt.shape)}")
                v = v.reshape(tgt.shape)
            tgt.data.copy_(v.to(self.device))

    # initialisers (distribution-faithful to the reference; TF's RNG stream cannot be matched)
    def glorot(self, shape, fan_in, fan_out):
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=self.gen, dtype=torch.float64) * 2 - 1).mul(lim).float()


# --------------------------------------------------------------------------------------------
# autograd wrappers (forward and backward are both library kernels)
# --------------------------------------------------------------------------------------------
def _as_time_major(x):
    """[B,T,C] -> contiguous [T,B,C].  The BLSTM stack hands its activations on as transposed VIEWS of
    time-major buffers, so between layers (and into the head) this is free; only a genuinely
    batch-major tensor (the stack's input) costs one transpose kernel."""
    xt = x.transpose(0, 1)
    if xt.is_contiguous():
        return xt
    return ops.transpose_01(x.contiguous())


class _BLSTMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kf, bf, kb, bb, precision):
        need_bwd = any(ctx.needs_input_grad)
        x_tm = _as_time_major(x)
        y_tm, saved = ops.blstm_fwd(x_tm, kf, bf, kb, bb, 1.0, precision, save_for_backward=need_bwd)
        if need_bwd:
            ctx.save_for_backward(x_tm, kf, kb, y_tm, saved)
        ctx.precision = precision
        return y_tm.transpose(0, 1)              # [B,T,2H] view of the time-major result

    @staticmethod
    def backward(ctx, dy):
        x_tm, kf, kb, y_tm, saved = ctx.saved_tensors
        dy_tm = _as_time_major(dy)
        dx, dkf, dbf, dkb, dbb = ops.blstm_bwd(x_tm, kf, kb, y_tm, dy_tm, saved, ctx.precision,
                                               need_dx=ctx.needs_input_grad[0])
        return (dx.transpose(0, 1) if dx is not None else None), dkf, dbf, dkb, dbb, None


class _DenseFn(torch.autograd.Function):
    """y[M,N] = x[M,K] @ W[K,N] + b   (tf.nn.conv1d with a [1,K,N] filter).
    swap=(B,T): the rows of x are time-major (t*B+b) and the rows of y batch-major (b*T+t) -- the
    [T,B,*] -> [B,T,*] hand-over between the BLSTM stack and the embedding reshape.  The small [M,K] activation is
    re-ordered (and kept for dW); re-ordering the wide [M,N] output rows in the GEMM epilogue instead scatters every
    128-row tile over 128 pages 10 MB apart and runs the head GEMM 3x slower (TLB misses).
    AMSS_PREC_BF16: x and W are converted to bf16 once; the copies feed the forward GEMM and are kept for the
    backward (dW = x^T dy reads x MN-major, dx = dy W^T reads W K-major: no transposes)."""

    @staticmethod
    def forward(ctx, x, W, b, precision, swap):
        if swap:
            Bq, Tq = swap
            x = ops.transpose_01(x.view(Tq, Bq, -1)).view(Bq * Tq, -1)
        ctx.precision, ctx.swap = precision, swap
        if precision == AMSS_PREC_FP32:
            ctx.save_for_backward(x, W)
            return ops.gemm(x, W, b, precision=precision)
        xb, Wb = ops.convert_bf16(x), ops.convert_bf16(W)
        ctx.save_for_backward(xb, Wb)
        ctx.bf16_operands = (xb, Wb)
        return ops.gemm_bf16(xb, False, Wb, True, x.shape[0], W.shape[1], W.shape[0], bias=b)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        swap = ctx.swap
        back_swap = (swap[1], swap[0]) if swap else None      # dy rows are batch-major, dx rows time-major
        dx = dW = db = None
        if ctx.precision == AMSS_PREC_FP32:
            if ctx.needs_input_grad[0]:
                dx = ops.gemm(dy, W, None, transb=True, precision=ctx.precision, out_swap=back_swap)
            if ctx.needs_input_grad[1]:
                dW = ops.gemm(x, dy, None, transa=True, precision=ctx.precision)
        else:
            M, N = dy.shape
            K = W.shape[0]
            dyb = ops.convert_bf16(dy)
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_bf16(dyb, False, W, False, M, K, N, out_swap=back_swap)
            if ctx.needs_input_grad[1]:
                dW = ops.gemm_bf16(x, True, dyb, True, K, N, M)
        if ctx.needs_input_grad[2]:
            db = ops.colsum(dy)
        return dx, dW, db, None, None


class _DenseNormFn(torch.autograd.Function):
    """V = l2_normalize(x W + b) over groups of E output columns on the tensor-core path: the normalisation runs in the
    GEMM epilogue (amss_gemm_bf16, norm_E), so the un-normalised [M,N] product is never written or re-read.
    Returns (V, inv_norm).  The backward is the generic one (dV -> dz -> dx, dW, db); dpcl_loss() replaces it by
    _HeadNormDPCLLossFn when the DPCL cost follows directly."""

    @staticmethod
    def forward(ctx, x, W, b, E, swap):
        if swap:                         # [T,B,C] -> [B*T,C] re-order fused with the bf16 conversion
            Bq, Tq = swap
            xb = ops.transpose_01_bf16(x.view(Tq, Bq, -1).contiguous())
        else:
            xb = ops.convert_bf16(x)
        Wb = ops.convert_bf16(W)
        V, inv = ops.gemm_bf16(xb, False, Wb, True, x.shape[0], W.shape[1], W.shape[0], bias=b, norm_E=E)
        ctx.save_for_backward(xb, Wb, V, inv)
        ctx.E, ctx.swap, ctx.bf16_operands = E, swap, (xb, Wb)
        ctx.mark_non_differentiable(inv)
        return V, inv

    @staticmethod
    def backward(ctx, dV, _dinv):
        xb, Wb, V, inv = ctx.saved_tensors
        swap = ctx.swap
        M, N = V.shape
        K = Wb.shape[0]
        dz = ops.l2norm_bwd(V, inv, dV.contiguous(), ctx.E)
        dzb = ops.convert_bf16(dz)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm_bf16(dzb, False, Wb, False, M, K, N, out_swap=(swap[1], swap[0]) if swap else None)
        if ctx.needs_input_grad[1]:
            dW = ops.gemm_bf16(xb, True, dzb, True, K, N, M)
        if ctx.needs_input_grad[2]:
            db = ops.colsum(dz)
        return dx, dW, db, None, None


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, E):
        v, inv = ops.l2norm_fwd(z.contiguous(), E)
        ctx.save_for_backward(v, inv)
        ctx.E = E
        ctx.mark_non_differentiable(inv)
        return v, inv

    @staticmethod
    def backward(ctx, dv, _dinv):
        v, inv = ctx.saved_tensors
        return ops.l2norm_bwd(v, inv, dv.contiguous(), ctx.E), None


class _DPCLLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, V, labels, S, precision):
        loss, ws = ops.dpcl_loss_fwd(V, labels, S, precision)
        ctx.save_for_backward(V, labels, ws)
        ctx.S = S
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        V, labels, ws = ctx.saved_tensors
        return ops.dpcl_loss_bwd(V, labels, ctx.S, dloss.reshape(1).contiguous(), ws), None, None, None


class _NormDPCLLossFn(torch.autograd.Function):
    """DPCL cost of V = l2_normalize(z) as ONE autograd node on z: the backward is a single kernel
    (affinity-loss gradient + normalisation Jacobian), dV never reaches HBM and v is read once."""

    @staticmethod
    def forward(ctx, z, V, inv, labels, S, precision):
        loss, ws = ops.dpcl_loss_fwd(V, labels, S, precision)
        ctx.save_for_backward(V, inv, labels, ws)
        ctx.S, ctx.zshape, ctx.precision = S, z.shape, precision
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        V, inv, labels, ws = ctx.saved_tensors
        dz = ops.dpcl_loss_bwd_normalized(V, labels, ctx.S, dloss.reshape(1).contiguous(), ws, inv, ctx.precision)
        return dz.view(ctx.zshape), None, None, None, None, None


class _WeightedDPCLLossFn(torch.autograd.Function):
    """DPCL cost with Y = weights * one_hot(labels) (--function_mask, models/network.py:381-389): fp32 kernels; one node
    on z when V = l2_normalize(z) (inv given), on V otherwise.  The weights are data (computed from the detached front
    output)."""

    @staticmethod
    def forward(ctx, z, V, inv, labels, weights, S):
        loss, ws = ops.dpcl_loss_weighted_fwd(V, labels, weights, S)
        ctx.save_for_backward(V, inv, labels, weights, ws)
        ctx.S, ctx.zshape = S, z.shape
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        V, inv, labels, weights, ws = ctx.saved_tensors
        dz = ops.dpcl_loss_weighted_bwd(V, labels, weights, ctx.S, dloss.reshape(1).contiguous(), ws, inv)
        return dz.view(ctx.zshape), None, None, None, None, None


class _HeadNormDPCLLossFn(torch.autograd.Function):
    """DPCL cost of l2_normalize(x W + b) as ONE autograd node on (x, W, b) for the tensor-core path: the backward
    writes dz once, in bf16, straight from the fused DPCL + normalisation kernel and feeds it to the two head GEMMs
    (dx = dz W^T, dW = x^T dz) and the bias column sum -- the fp32 [B*T, F*E] gradient never exists."""

    @staticmethod
    def forward(ctx, x, W, b, xb, Wb, V, inv, labels, S, swap):
        loss, ws = ops.dpcl_loss_fwd(V, labels, S, ops.AMSS_PREC_BF16)
        ctx.save_for_backward(xb, Wb, V, inv, labels, ws)
        ctx.S, ctx.swap, ctx.K, ctx.N = S, swap, W.shape[0], W.shape[1]
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        xb, Wb, V, inv, labels, ws = ctx.saved_tensors
        swap, K, N = ctx.swap, ctx.K, ctx.N
        dzb = ops.dpcl_loss_bwd_normalized_bf16(V, labels, ctx.S, dloss.reshape(1).contiguous(), ws, inv)
        M = dzb.numel() // N
        dzb = dzb.view(M, N)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm_bf16(dzb, False, Wb, False, M, K, N, out_swap=(swap[1], swap[0]) if swap else None)
        if ctx.needs_input_grad[1]:
            dW = ops.gemm_bf16(xb, True, dzb, True, K, N, M)
        if ctx.needs_input_grad[2]:
            db = ops.colsum_bf16(dzb)
        return dx, dW, db, None, None, None, None, None, None, None


class _L41LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, labels, spk, weights):
        ctx.save_for_backward(emb, labels, spk, weights)
        return ops.l41_loss_fwd(emb, labels, spk, weights).view(())

    @staticmethod
    def backward(ctx, dloss):
        emb, labels, spk, weights = ctx.saved_tensors
        demb, dspk = ops.l41_loss_bwd(emb, labels, spk, dloss.reshape(1).contiguous(), weights)
        return demb, None, dspk, None


class _MakeFilterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, window, bases):
        ctx.save_for_backward(window, bases)
        return ops.make_filter(window, bases)

    @staticmethod
    def backward(ctx, dfilt):
        window, bases = ctx.saved_tensors
        return ops.make_filter_bwd(window, bases, dfilt.contiguous())


class _AnalysisFn(torch.autograd.Function):
    """conv2d SAME stride 1 + max_pool_with_argmax, fused (reference models/adapt.py:115-117).
    Gradient flows to the filter only (the waveform is data)."""

    @staticmethod
    def forward(ctx, x, filt, pool, hop, precision, batch=None):
        if batch is not None:       # x = [B mixtures ; B*S sources]: lets the library use the linear-mixture fast path
            y, am = ops.filterbank_analysis_mix(x, filt, batch[0], batch[1], pool, hop, precision)
        else:
            y, am = ops.filterbank_analysis(x, filt, pool, hop, ops.AMSS_POOL_MAX, precision)
        ctx.save_for_backward(x, am)
        ctx.W = filt.shape[0]
        ctx.mark_non_differentiable(am)
        return y, am

    @staticmethod
    def backward(ctx, dy, _dam):
        x, am = ctx.saved_tensors
        return None, ops.filterbank_analysis_bwd(x, dy.contiguous(), am, ctx.W), None, None, None, None


def strided_positions(Bt, L, W, N, hop, device):
    """The strided front end (tf.nn.conv2d, strides [1,1,hop,1], SAME; models/adapt.py:121-122) computes
    y[r,tp,n] = sum_k x[tp*hop + k - pl_s] filt[k,n] with TF's SAME split pl_s = max((ceil(L/hop)-1)*hop + W - L, 0) // 2,
    i.e. the stride-1 response at the FIXED position pos = tp*hop - pl_s + (W-1)//2.  Written as a per-sample flat index
    pos*N + n -- the convention of max_pool_with_argmax -- these positions let the sparse kernels of the max-pool path
    (filter gradient through the arg-max, unpool + transposed-conv synthesis and their gradients) serve the strided
    mode unchanged: -> (int64 [Bt,Tp,N], offset c = pos - tp*hop)."""
    Tp = -(-L // hop)
    pl_s = max((Tp - 1) * hop + W - L, 0) // 2
    c = (W - 1) // 2 - pl_s
    pos = torch.arange(Tp, device=device, dtype=torch.int64) * hop + c
    am = pos.view(1, Tp, 1) * N + torch.arange(N, device=device, dtype=torch.int64).view(1, 1, N)
    return am.expand(Bt, Tp, N).contiguous(), c


class _AnalysisStridedFn(torch.autograd.Function):
    """Strided front end with the filter gradient (sparse kernel at the fixed positions, see strided_positions)."""

    @staticmethod
    def forward(ctx, x, filt, hop):
        y, _ = ops.filterbank_analysis(x, filt, hop, hop, ops.AMSS_POOL_STRIDE, AMSS_PREC_FP32)
        am, _ = strided_positions(x.shape[0], x.shape[1], filt.shape[0], filt.shape[1], hop, x.device)
        ctx.save_for_backward(x, am)
        ctx.W = filt.shape[0]
        ctx.mark_non_differentiable(am)
        return y, am

    @staticmethod
    def backward(ctx, dy, _dam):
        x, am = ctx.saved_tensors
        return None, ops.filterbank_analysis_bwd(x, dy.contiguous(), am, ctx.W), None


def analysis_strided(x, filt, hop):
    return _AnalysisStridedFn.apply(x, filt, hop)


class _AnalysisAvgFn(torch.autograd.Function):
    """Average-pool front end (conv2d SAME stride 1 + average_pooling2d(pool, stride pool), models/adapt.py:118-120) with the
    filter gradient: y[r,tp,n] = sum_k xs[r, tp*P + k - pl] filt[k,n] with the box-filtered signal
    xs[u] = (1/P) sum_{j<P} x[u + j] (u from -(P-1) on), i.e. the sparse filter-gradient kernel at the fixed positions tp*P on xs."""

    @staticmethod
    def forward(ctx, x, filt, pool):
        y, _ = ops.filterbank_analysis(x, filt, pool, pool, ops.AMSS_POOL_AVG, AMSS_PREC_FP32)
        ctx.save_for_backward(x)
        ctx.W, ctx.N, ctx.pool, ctx.Tp = filt.shape[0], filt.shape[1], pool, y.shape[1]
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        P, N, Tp = ctx.pool, ctx.N, ctx.Tp
        # xs is needed for u in [-(P-1), L) (the window of a frame's first taps reaches before the signal): it is stored shifted
        # by P - 1, xs_ext[i] = (1/P) sum_{j<P} x[i - j], and the positions move with it
        xs = ops.box_sum(x.contiguous(), P, dir=-1, scale=1.0 / P, len_out=x.shape[1] + P - 1)
        pos = torch.arange(Tp, device=x.device, dtype=torch.int64) * P + (P - 1)
        am = (pos.view(1, Tp, 1) * N + torch.arange(N, device=x.device, dtype=torch.int64).view(1, 1, N))
        am = am.expand(x.shape[0], Tp, N).contiguous()
        return None, ops.filterbank_analysis_bwd(xs, dy.contiguous(), am, ctx.W), None


def analysis_avg(x, filt, pool):
    return _AnalysisAvgFn.apply(x, filt, pool)


class _BoxFilterFn(torch.autograd.Function):
    """wbox[k', n] = sum_{j<P} filt[k' - j, n], k' < W + P - 1 (columns of a [W, N] bank); adjoint: sum_{j<P} dwbox[k + j, n]."""

    @staticmethod
    def forward(ctx, filt, P):
        ctx.P, ctx.W = P, filt.shape[0]
        return ops.box_sum(filt.contiguous(), P, dir=-1, len_out=filt.shape[0] + P - 1, axis=0)

    @staticmethod
    def backward(ctx, dwbox):
        return ops.box_sum(dwbox.contiguous(), ctx.P, dir=1, len_out=ctx.W, axis=0), None


def synthesis_avg(vals, filt2, B, S, L, pool):
    """Average-pool back end (UpSampling2D((1, pool)) + conv2d_transpose SAME stride 1, models/adapt.py:224-243): every pooled
    value is an atom at the fixed position tp*pool of the BOX-FILTERED bank wbox (W + pool - 1 taps); the positions are
    shifted by the difference of the two banks' left paddings so that the sparse overlap-add indexes wbox as the reference's
    transposed convolution indexes filt2.  Gradients flow to vals and (through the box filter's adjoint) to filt2."""
    W, N = filt2.shape
    Tp = vals.shape[1]
    wbox = _BoxFilterFn.apply(filt2, pool)
    c = (W + pool - 2) // 2 - (W - 1) // 2
    pos = torch.arange(Tp, device=vals.device, dtype=torch.int64) * pool + c
    am = (pos.view(1, Tp, 1) * N + torch.arange(N, device=vals.device, dtype=torch.int64).view(1, 1, N)).expand(B, Tp, N).contiguous()
    return _SynthesisFn.apply(vals, am, wbox, B, S, L, c + 1, pool)


class _SynthesisFn(torch.autograd.Function):
    """unpool + conv2d_transpose fused as a sparse overlap-add (reference models/adapt.py:205-252)."""

    @staticmethod
    def forward(ctx, vals, argmax_mix, filt2, B, S, L, pool, hop):
        ctx.save_for_backward(vals, argmax_mix, filt2)
        ctx.dims = (B, S)
        return ops.filterbank_synthesis(vals, argmax_mix, filt2, B, S, L, pool, hop)

    @staticmethod
    def backward(ctx, dout):
        vals, am, filt2 = ctx.saved_tensors
        B, S = ctx.dims
        dvals, dfilt2 = ops.filterbank_synthesis_bwd(dout.contiguous(), vals, am, filt2, B, S,
                                                     ctx.needs_input_grad[0], ctx.needs_input_grad[2])
        return dvals, None, dfilt2, None, None, None, None, None


class _WaveLossFn(torch.autograd.Function):
    """Per-(b,s) waveform statistics of the Adapt pre-training cost (models/adapt.py:323-330,
    models/network.py:207-211), one fused reduction kernel: stats[r] = (<t,t>, <a,a>, <t,a>, <t-a,t-a>).
    Gradient flows to the approximation a only (the targets are data):
        d<a,a> = 2a, d<t,a> = t, d<t-a,t-a> = 2(a - t)."""

    @staticmethod
    def forward(ctx, target, approx):
        target, approx = target.contiguous(), approx.contiguous()
        ctx.save_for_backward(target, approx)
        return ops.wave_stats(target, approx)

    @staticmethod
    def backward(ctx, dstats):
        target, approx = ctx.saved_tensors
        return None, ops.wave_stats_bwd(target, approx, dstats.contiguous())


def wave_stats(target, approx):
    return _WaveLossFn.apply(target, approx)


class _AdaptCostFn(torch.autograd.Function):
    """The scalar tail of Adapt.cost (pre-training branch) as ONE node: waveform statistics, the three front-output terms
    and the two filter banks -> cost, with (l2, sdr, sdr_improvement) as by-products; forward and every derivative come from
    amss_adapt_cost_fwd (the tensor-op version was ~90 launches of 2-4 us on a handful of numbers)."""

    @staticmethod
    def forward(ctx, st, terms, filt, filt2, sm, B, S, loss_kind, beta, lam, ov, nn):
        regsq = ops.sumsq([filt, filt2]) if lam != 0.0 else torch.zeros(1, dtype=st.dtype, device=st.device)
        out4, dst, dterms, dreg = ops.adapt_cost_fwd(st.contiguous(), sm, terms.contiguous(), regsq, B, S, loss_kind, beta, lam, ov, nn)
        ctx.save_for_backward(dst, dterms, dreg, filt, filt2)
        ctx.lam = lam
        aux = out4[1:].clone()
        ctx.mark_non_differentiable(aux)
        return out4[0].clone(), aux

    @staticmethod
    def backward(ctx, dcost, _daux):
        dst, dterms, dreg, filt, filt2 = ctx.saved_tensors
        g = dcost.reshape(())
        dfilt = dfilt2 = None
        if ctx.lam != 0.0:
            c = dreg[0] * g
            dfilt = filt * c if ctx.needs_input_grad[2] else None
            dfilt2 = filt2 * c if ctx.needs_input_grad[3] else None
        return dst * g, dterms * g, dfilt, dfilt2, None, None, None, None, None, None, None, None


def adapt_cost(st, terms, filt, filt2, sm, B, S, loss, beta, lam, ov, nn):
    """-> (cost, aux[3] = l2, sdr, sdr_improvement)."""
    kind = 0 if loss == "l2" else (1 if loss == "sdr" else 2)
    return _AdaptCostFn.apply(st, terms, filt, filt2, sm, B, S, kind, float(beta), float(lam), float(ov), float(nn))


class _ISTFTMaskedFn(torch.autograd.Function):
    """Separator.postprocessing with soft masks (network.py:584-607) as a differentiable node: (masks * X) ->
    inverse_stft.  The gradient reaches the masks (the enhance layer's softmax output); the mixture STFT is data."""

    @staticmethod
    def forward(ctx, spec, masks, S, frame, hop):
        ctx.save_for_backward(spec)
        ctx.cfg = (S, frame, hop)
        return ops.istft_masked(spec, S, frame, hop, masks=masks.contiguous())

    @staticmethod
    def backward(ctx, dout):
        spec, = ctx.saved_tensors
        S, frame, hop = ctx.cfg
        return None, ops.istft_masked_bwd(spec, dout.contiguous(), S, frame, hop), None, None, None


def istft_masked(spec, masks, S, frame, hop):
    return _ISTFTMaskedFn.apply(spec, masks, S, frame, hop)


class _AdaptTermsFn(torch.autograd.Function):
    """The reduction / elementwise terms of the Adapt pre-training graph over the front output, one fused kernel each
    way (amss_adapt_terms_fwd / _bwd): y -> (separator output, p_hat, [sparse_constraint, overlapping, nonneg])."""

    @staticmethod
    def forward(ctx, y, B, S, rho, separation, want_sep):
        y = y.contiguous()
        sep, p_hat, terms = ops.adapt_terms_fwd(y, B, S, rho, separation, want_sep)
        ctx.save_for_backward(y, p_hat)
        ctx.cfg = (B, S, rho, separation)
        ctx.mark_non_differentiable(p_hat)
        if sep is None:
            sep = y.new_zeros(0)
        return sep, p_hat, terms

    @staticmethod
    def backward(ctx, dsep, _dp, dterms):
        y, p_hat = ctx.saved_tensors
        B, S, rho, separation = ctx.cfg
        dsep = dsep.contiguous() if dsep is not None and dsep.numel() else None
        dterms = dterms.contiguous() if dterms is not None else torch.zeros(3, dtype=y.dtype, device=y.device)
        return ops.adapt_terms_bwd(y, p_hat, dsep, dterms, B, S, rho, separation), None, None, None, None, None


def adapt_terms(y, B, S, rho, separation, want_sep=True):
    return _AdaptTermsFn.apply(y, B, S, rho, separation, want_sep)


class _EnhanceCostFn(torch.autograd.Function):
    """enhance_cost (models/network.py:662-693) of the enhance Conv1D output, fused with the softmax / tanh over the
    sources and the multiplication with X_input (:640-656): one kernel builds the [B,S,S] distance table, the best of
    the S! permutations is picked from it (S*S scalars per mixture), one kernel writes d cost / d logits."""

    @staticmethod
    def forward(ctx, logits, X_input, X_non_mix, nonlinearity):
        import itertools
        logits, X_input, X_non_mix = logits.contiguous(), X_input.contiguous(), X_non_mix.contiguous()
        B, S, TF = logits.shape
        table, _ = ops.enhance_cost_table(logits, X_input, X_non_mix, nonlinearity)
        perms = list(itertools.permutations(range(S)))
        pt = torch.tensor(perms, device=logits.device)                               # [P,S]
        costs = table[:, torch.arange(S, device=logits.device).unsqueeze(0), pt].sum(-1)     # [B,P]
        best = costs.argmin(1)
        ctx.save_for_backward(logits, X_input, X_non_mix, pt[best].to(torch.int32).contiguous())
        ctx.nonlinearity = nonlinearity
        return costs.gather(1, best.unsqueeze(1)).mean()

    @staticmethod
    def backward(ctx, dcost):
        logits, X_input, X_non_mix, perm = ctx.saved_tensors
        B = logits.shape[0]
        dcost_b = (dcost / B).reshape(1).expand(B).contiguous()
        return ops.enhance_cost_bwd(logits, X_input, X_non_mix, perm, dcost_b, ctx.nonlinearity), None, None, None


def enhance_cost_fused(logits, X_input, X_non_mix, nonlinearity="softmax"):
    return _EnhanceCostFn.apply(logits, X_input, X_non_mix, nonlinearity)


class _PairDotsFn(torch.autograd.Function):
    """G[b, b'] = <t[b], a[b']> over L (library GEMM t a^T); gradient to a only: da = dG^T t."""

    @staticmethod
    def forward(ctx, t, a):
        ctx.save_for_backward(t)
        return ops.gemm(t, a, None, transb=True)

    @staticmethod
    def backward(ctx, dG):
        t, = ctx.saved_tensors
        return None, ops.gemm(dG.contiguous(), t, None, transa=True)


def pair_dots(t, a):
    return _PairDotsFn.apply(t, a)


def pit_wave_l2(x_non_mix, est, reduce="mean"):
    """cost_finetuning (models/network.py:697-723, models/adapt.py:404-431): 0.5 * sum_L (x_s - xhat_perm(s))^2, mean
    over the S sources, min over the S! permutations, mean over the batch.  The S*S pairwise squared distances come
    from S launches of the fused waveform-statistics kernel (pair s with estimate (s+k) mod S); the permutation
    min/mean runs on the [B,S,S] table.  Gradient flows to `est` through the kernel's autograd wrapper."""
    import itertools
    B, S, Lw = x_non_mix.shape
    tgt = x_non_mix.reshape(B * S, Lw).contiguous()
    d = []
    for k in range(S):
        idx = [(sidx + k) % S for sidx in range(S)]
        e = (est if k == 0 else est[:, idx]).reshape(B * S, Lw).contiguous()
        d.append(0.5 * wave_stats(tgt, e)[:, 3].reshape(B, S))
    D = torch.stack(d, 2)                                    # D[b, s, k] = 0.5 * |x[b,s] - est[b,(s+k)%S]|^2
    red = (lambda t: t.mean(1)) if reduce == "mean" else (lambda t: t.sum(1))          # over the S sources
    costs = [red(torch.stack([D[:, sidx, (perm[sidx] - sidx) % S] for sidx in range(S)], 1))
             for perm in itertools.permutations(range(S))]
    return torch.stack(costs, 1).min(1).values.mean()


def blstm(x, kf, bf, kb, bb, precision=AMSS_PREC_FP32):
    return _BLSTMFn.apply(x, kf, bf, kb, bb, precision)


def _carry(src, dst):
    """Views made by the layer protocol (Conv1D's [B,T,N] view, Reshape) keep the side-channel that lets dpcl_loss()
    fuse the head's backward with the loss (see _HeadNormDPCLLossFn)."""
    for attr in ("_amss_dense", "_amss_head"):
        h = getattr(src, attr, None)
        if h is not None:
            setattr(dst, attr, h)
    return dst


def dense(x, W, b, precision=AMSS_PREC_FP32, swap=None):
    y = _DenseFn.apply(x, W, b, precision, swap)
    if precision != AMSS_PREC_FP32 and y.requires_grad and b is not None:
        fn = y.grad_fn
        ops_b = getattr(fn, "bf16_operands", None)
        if ops_b is not None:
            y._amss_dense = (x, W, b, ops_b[0], ops_b[1], swap)
    return y


def dense_normalized(x, W, b, E, swap=None):
    """l2_normalize(dense(x)) with the normalisation fused into the GEMM epilogue (tensor-core path only)."""
    V, inv = _DenseNormFn.apply(x, W, b, E, swap)
    if V.requires_grad:
        xb, Wb = V.grad_fn.bf16_operands
        V._amss_head = (x, W, b, xb, Wb, swap, inv)
    return V


def l2_normalize(z, E):
    v, inv = _L2NormFn.apply(z, E)
    if z.requires_grad:
        v._amss_prenorm = (z, inv)          # lets dpcl_loss() fuse the two backward passes (see _NormDPCLLossFn)
    return v


def dpcl_loss(V, labels, S, prenorm=None, precision=AMSS_PREC_FP32, head=None, weights=None):
    """prenorm = (z, inv_norm) as stashed by l2_normalize() on its output: the loss becomes one autograd node
    on z with a fused backward (DPCL gradient + normalisation Jacobian).  On the tensor-core path, when z is the
    output of a dense layer, the node moves one step further up (onto the layer's x, W, b): _HeadNormDPCLLossFn.
    weights [B,TF] (--function_mask): the weighted label matrix, fp32 kernels whatever the precision."""
    if weights is not None:
        weights = weights.reshape(labels.shape).contiguous()
        # (a V that came out of dense_normalized() keeps its generic backward: dV -> dz -> dx, dW, db)
        if prenorm is not None and head is None:
            z, inv = prenorm
            if z.numel() == V.numel():
                return _WeightedDPCLLossFn.apply(z, V.detach().contiguous(), inv, labels, weights, S)
        Vc = V.contiguous()
        return _WeightedDPCLLossFn.apply(Vc, Vc, None, labels, weights, S)
    if head is not None and precision != AMSS_PREC_FP32 and S <= 4:
        # V came out of dense_normalized(): (x, W, b, bf16 copies, swap, inv_norm)
        x, W, b, xb, Wb, swap, inv = head
        return _HeadNormDPCLLossFn.apply(x, W, b, xb, Wb, V.detach(), inv, labels, S, swap)
    if prenorm is not None:
        z, inv = prenorm
        head = getattr(z, "_amss_dense", None)
        E = V.shape[-1]
        if head is not None and precision != AMSS_PREC_FP32 and E % 8 == 0 and 8 <= E <= 64 and S <= 4:
            x, W, b, xb, Wb, swap = head
            if W.shape[1] % 2 == 0 and z.numel() == V.numel():
                return _HeadNormDPCLLossFn.apply(x, W, b, xb, Wb, V.detach(), inv, labels, S, swap)
        return _NormDPCLLossFn.apply(z, V.detach(), inv, labels, S, precision)
    return _DPCLLossFn.apply(V.contiguous(), labels, S, precision)


def l41_loss(emb, labels, spk, weights=None):
    return _L41LossFn.apply(emb.contiguous(), labels, spk, weights)


def make_filter(window, bases):
    return _MakeFilterFn.apply(window, bases)


def analysis(x, filt, pool, hop, precision=AMSS_PREC_FP32, batch=None):
    return _AnalysisFn.apply(x, filt, pool, hop, precision, batch)


def synthesis(vals, argmax_mix, filt2, B, S, L, pool, hop):
    return _SynthesisFn.apply(vals, argmax_mix, filt2, B, S, L, pool, hop)


# --------------------------------------------------------------------------------------------
# the reference's layer classes (utils/ops.py)
# --------------------------------------------------------------------------------------------
def f_props(layers, x):
    """utils/ops.py:82-85."""
    for layer in layers:
        x = layer.f_prop(x)
    return x


class BLSTM:
    """utils/ops.py:358-383.  hid_dim is the concatenated width: each direction has hid_dim//2 cells.
    Variables: <scope>/{forward,backward}_<name>/rnn/basic_lstm_cell/{kernel,bias}."""

    def __init__(self, hid_dim, name, drop_val=0.0, *, store, scope, in_dim, precision=AMSS_PREC_FP32):
        if drop_val != 0.0:
            raise NotImplementedError("recurrent dropout > 0 is outside the hot path (default 0.0, utils/trainer.py:77-78)")
        self.hid_dim, self.name, self.precision = hid_dim, name, precision
        H = hid_dim // 2
        self.keys = []
        for d in ("forward", "backward"):
            base = f"{scope}/{d}_{name}/rnn/basic_lstm_cell"
            store.register(base + "/kernel", store.glorot((in_dim + H, 4 * H), in_dim + H, 4 * H))
            store.register(base + "/bias", torch.zeros(4 * H))
            self.keys += [base + "/kernel", base + "/bias"]
        self.store = store

    def f_prop(self, x):
        kf, bf, kb, bb = (self.store[k] for k in self.keys)
        return blstm(x, kf, bf, kb, bb, self.precision)


class Conv1D:
    """utils/ops.py:486-503 with filter_shape [1, in, out]: a per-frame dense layer.
    reference_scale=True reproduces the reference's (very wide) uniform init range."""

    def __init__(self, filter_shape, *, store, scope, name="Conv1D", precision=AMSS_PREC_FP32, reference_scale=False):
        assert filter_shape[0] == 1, "only kernel size 1 is on the hot path"
        _, cin, cout = filter_shape
        if reference_scale:
            fan = math.sqrt(2.0 / float(cin + cout))
            lim = math.sqrt(2.0 / fan)
        else:
            lim = math.sqrt(6.0 / (cin + cout))
        W = (torch.rand((cin, cout), generator=store.gen, dtype=torch.float64) * 2 - 1).mul(lim).float()
        store.register(f"{scope}/W", W)
        store.register(f"{scope}/b", torch.zeros(cout))
        self.store, self.scope, self.precision = store, scope, precision

    def f_prop(self, x):
        B, Tt, C = x.shape
        W, b = self.store[f"{self.scope}/W"], self.store[f"{self.scope}/b"]
        xt = x.transpose(0, 1)
        if xt.is_contiguous() and not x.is_contiguous():
            # time-major activations straight from the BLSTM stack: the GEMM epilogue remaps the rows
            y = dense(xt.reshape(Tt * B, C), W, b, self.precision, swap=(B, Tt))
        else:
            y = dense(x.reshape(B * Tt, C), W, b, self.precision)
        return _carry(y, y.view(B, Tt, -1))

    def f_prop_normalized(self, x, E):
        """l2_normalize(f_prop(x)) over groups of E output channels (the Reshape + Normalize that follow the head in
        models/dpcl.py:30-37), fused into the GEMM epilogue.  Tensor-core path only."""
        B, Tt, C = x.shape
        W, b = self.store[f"{self.scope}/W"], self.store[f"{self.scope}/b"]
        xt = x.transpose(0, 1)
        if xt.is_contiguous() and not x.is_contiguous():
            V = dense_normalized(xt.reshape(Tt * B, C), W, b, E, swap=(B, Tt))
        else:
            V = dense_normalized(x.reshape(B * Tt, C), W, b, E)
        return _carry(V, V.view(B, Tt, -1))


class Reshape:
    """utils/ops.py:310-316."""

    def __init__(self, shape, name="Reshape"):
        self.shape, self.name = shape, name

    def f_prop(self, x):
        return _carry(x, x.reshape(self.shape))


class Normalize:
    """utils/ops.py:318-324: tf.nn.l2_normalize over the (last) axis."""

    def __init__(self, axis, name="Normalize"):
        self.axis, self.name = axis, name

    def f_prop(self, x):
        assert self.axis in (-1, x.dim() - 1), "l2_normalize is implemented over the innermost axis"
        return l2_normalize(x, x.shape[-1])
