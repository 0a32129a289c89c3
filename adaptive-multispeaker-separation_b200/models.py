"""Host-side mirror of the reference's model classes for the hot path: Adapt (models/adapt.py),
Separator / DPCL / L41Model (models/network.py, models/dpcl.py, models/L41.py) and KMeans
(models/Kmeans_2.py).  Constructors take the reference's flat kwargs dict (CLI flag names,
utils/trainer.py:17-166); methods keep the reference's property names (front, separator, back,
cost, preprocessing, prediction, separate, postprocessing).

The reference builds a TF graph once and runs it with sess.run; here the same nodes are plain
methods that launch the library's kernels on torch CUDA tensors (eager, taped by torch.autograd).
"""
import itertools
import math

import numpy as np
import torch

from . import layers as L
from . import ops
from ._lib import AMSS_PREC_FP32, AMSS_PREC_BF16, AmssError

DEFAULTS = dict(
    # common (utils/trainer.py:17-48)
    chunk_size=20480, nb_speakers=2, batch_size=64, epochs=10, learning_rate=0.1, optimizer="Adam",
    decay_epoch=50, gradient_norm_clip=0.0, validation_step=1000, dataset_normalize=False,
    # separator (utils/trainer.py:57-109)
    normalize_separator="None", abs_input=False, pre_func="None", silence_mask_db=0, nb_layers=3, layer_size=600,
    embedding_size=40, no_normalize=True, recurrent_dropout=0.0, nb_tries=10, nb_steps=10, beta_kmeans=None,
    threshold=2.0, with_silence=False, end_assign=False, silence_loss=False, threshold_silence_loss=2.0,
    function_mask="None", sampling=None, ns_rate=0.1, ns_method="random", add_dilated=False,
    tot_speakers=251,
    # enhance (utils/trainer.py:122-132)
    normalize_enhance=False, nb_layers_enhance=3, layer_size_enhance=600, nonlinearity="softmax",
    recurrent_dropout_enhance=0.0,
    # adapt (utils/trainer.py:134-166)
    filters=512, max_pool=512, with_max_pool=False, with_average_pool=False, regularization=1e-4, beta=1e-2,
    sparsity=0.01, overlap_coef=0.001, overlap_value=0.1, non_negativity=0.0, loss="sdr", separation="perfect",
    pretraining=True,
    # this implementation
    precision="fp32", seed=42, reference_init=False,
)


def _prec(v):
    if v in (AMSS_PREC_FP32, AMSS_PREC_BF16):
        return v
    return {"fp32": AMSS_PREC_FP32, "bf16": AMSS_PREC_BF16}[str(v)]


class Network:
    """models/network.py:12-63 -- holds the flat kwargs and the shared parameter store."""

    def __init__(self, store=None, **kwargs):
        unknown = set(kwargs) - set(DEFAULTS) - {"window_size", "hop_size", "type", "pipeline", "mix", "non_mix",
                                                  "ind", "model_folder", "sex", "train", "dataset", "men", "women",
                                                  "no_random_picking", "mask_a", "mask_b"}
        if unknown:
            raise KeyError(f"unknown arguments: {sorted(unknown)}")
        self.args = dict(DEFAULTS)
        self.args.update(kwargs)
        self.S = self.args["nb_speakers"]
        self.precision = _prec(self.args["precision"])
        self.store = store if store is not None else L.ParamStore(seed=self.args["seed"])
        self._owns_store = store is None

    def finalize(self):
        if self.store.flat is None:
            self.store.finalize()
        return self

    # ---- model folder: `params` JSON + variables under the reference's names -------------------
    # models/network.py:124-129 writes the flat kwargs (minus the input tensors) to <folder>/params, :223-226 saves the
    # variables, :291-306 (`load`) rebuilds a model from that file letting a fixed list of keys be overridden.  TF's
    # checkpoint bundle cannot be written or read without TensorFlow; the variables go to <folder>/model.npz keyed by
    # the same names a TF checkpoint of the reference graph uses, so a converter is a rename-free
    # dump of `tf.train.load_checkpoint(...).get_tensor(name)` on a machine that has TF 1.x.
    KEYS_TO_UPDATE = ("learning_rate", "epochs", "batch_size", "chunk_size", "nb_speakers", "regularization", "overlap_coef",
                      "loss", "beta", "model_folder", "type", "pretraining", "with_silence", "end_assign", "beta_kmeans",
                      "nb_tries", "nb_steps", "threshold", "optimizer", "men", "women", "recurrent_dropout",
                      "recurrent_dropout_enhance")

    def save(self, folder, tf_checkpoint=False, step=0):
        """<folder>/params (JSON) + <folder>/model.npz; tf_checkpoint=True also writes <folder>/model-<step>.index /
        .data-00000-of-00001 in TensorFlow's tensor-bundle format (tf_bundle.py) with Conv1D filters in the
        reference's [1, in, out] shape (utils/ops.py:486-494) plus the `checkpoint` state file tf.train.latest_checkpoint
        reads (models/network.py:254-262)."""
        import json
        import os
        os.makedirs(folder, exist_ok=True)
        args = {k: v for k, v in self.args.items() if k not in ("mix", "non_mix", "ind")}
        with open(os.path.join(folder, "params"), "w") as f:
            json.dump(args, f)
        self.finalize()
        sd = {k: v.numpy() for k, v in self.store.state_dict().items()}
        np.savez(os.path.join(folder, "model.npz"), **sd)
        if tf_checkpoint:
            from . import tf_bundle
            tfv = {k: (v[None] if k.endswith("/W") and v.ndim == 2 else v) for k, v in sd.items()}
            tf_bundle.save_checkpoint(os.path.join(folder, f"model-{step}"), tfv)
            with open(os.path.join(folder, "checkpoint"), "w") as f:
                f.write(f'model_checkpoint_path: "model-{step}"\nall_model_checkpoint_paths: "model-{step}"\n')
        return folder

    @classmethod
    def load(cls, path, modified_args=None):
        """models/network.py:291-306: the stored kwargs, with `modified_args` applied for the keys in KEYS_TO_UPDATE and
        for keys the stored file does not have.  Variables are restored separately (restore_model), as in the reference."""
        import json
        import os
        modified_args = dict(modified_args or {})
        with open(os.path.join(path, "params")) as f:
            args = json.load(f)
        upd = {k: modified_args[k] for k in cls.KEYS_TO_UPDATE if k in modified_args}
        upd.update({k: v for k, v in modified_args.items() if k not in args})
        args.update(upd)
        return cls(**args)

    def restore_model(self, path, strict=False):
        """models/network.py:254-262: load every stored variable this model also has (by name).  `path` is a model folder
        holding model.npz, or a folder / prefix of a TensorFlow checkpoint written by the reference's Saver
        (`checkpoint` state file -> latest `model-<step>`; read without TensorFlow by tf_bundle.py; Conv1D filters
        [1,in,out] are reshaped by load_state_dict)."""
        import os
        import re
        self.finalize()
        npz = os.path.join(path, "model.npz")
        if os.path.isdir(path) and os.path.exists(npz):
            with np.load(npz) as z:
                self.store.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=strict)
            return self
        from . import tf_bundle
        prefix = path
        if os.path.isdir(path):                                  # tf.train.latest_checkpoint(path)
            state = os.path.join(path, "checkpoint")
            if not os.path.exists(state):
                raise FileNotFoundError(f"{path}: neither model.npz nor a TensorFlow `checkpoint` state file")
            m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', open(state).read())
            prefix = m.group(1) if os.path.isabs(m.group(1)) else os.path.join(path, m.group(1))
        names = set(self.store.names())
        tensors = tf_bundle.load_checkpoint(prefix, names=names)
        # TF appends ":0" to tensor names but stores variables under their op names; Adam/AMSGrad slot variables
        # (".../AMSGrad", ".../AMSGrad_1") and global_epoch are simply not part of this model's names
        self.store.load_state_dict({k: torch.from_numpy(v) for k, v in tensors.items()}, strict=strict)
        return self

    # the reference's freeze_all_with(prefix) (models/network.py:276-289)
    def freeze_all_with(self, prefix):
        self.store.set_trainable(lambda n, p=prefix: not n.startswith(p) and self.store[n].requires_grad)

    def freeze_all_except(self, *prefixes):
        self.store.set_trainable(lambda n: any(n.startswith(p) for p in prefixes))


# =================================================================================================
# Adaptive front / back end                                                    models/adapt.py
# =================================================================================================
class Adapt(Network):
    def __init__(self, store=None, **kwargs):
        kwargs.setdefault("window_size", 1024)
        kwargs.setdefault("hop_size", 256)
        super().__init__(store, **kwargs)
        a = self.args
        self.N, self.max_pool_value, self.window = a["filters"], a["max_pool"], a["window_size"]
        self.hop_size, self.pretraining = a["hop_size"], a["pretraining"]
        self.l, self.beta, self.p = a["regularization"], a["beta"], a["sparsity"]
        self.overlap_coef, self.loss, self.separation = a["overlap_coef"], a["loss"], a["separation"]
        self.with_max_pool, self.with_average_pool = a["with_max_pool"], a["with_average_pool"]
        self.non_negativity = a["non_negativity"]
        st, W, N = self.store, self.window, self.N
        # xavier_initializer_conv2d variables (adapt.py:104-105, :232-233)
        st.register("front/window/w", st.glorot((W,), W, 1))
        st.register("front/bases/bases", st.glorot((W, N), W, N))
        st.register("back/window/value", st.glorot((W,), W, 1))
        st.register("back/bases/value", st.glorot((W, N), W, N))
        self.sepNet = None
        if self._owns_store and self.pretraining:
            self.finalize()

    @property
    def pool_mode(self):
        if self.with_max_pool:
            return ops.AMSS_POOL_MAX
        return ops.AMSS_POOL_AVG if self.with_average_pool else ops.AMSS_POOL_STRIDE

    def conv_filter(self, scope="front"):
        if scope == "front":
            return L.make_filter(self.store["front/window/w"], self.store["front/bases/bases"])
        return L.make_filter(self.store["back/window/value"], self.store["back/bases/value"])

    # adapt.py:41-48 + 95-134
    def front(self, x_mix, x_non_mix):
        """-> (y [B(S+1),Tp,N], argmax int64 or None).  Mixture rows first, then the B*S sources."""
        B, S, Lw = x_non_mix.shape
        x = torch.cat([x_mix, x_non_mix.reshape(B * S, Lw)], 0).contiguous()
        filt = self.conv_filter("front")
        if self.with_max_pool:
            y, am = L.analysis(x, filt, self.max_pool_value, self.hop_size, self.precision, batch=(B, S))
        elif not self.with_average_pool:
            # the reference's DEFAULT front end: strided convolution (adapt.py:121-122).  `am` holds the fixed positions
            # tp*hop + c in the arg-max convention, so back() and both filter gradients run on the sparse kernels
            y, am = L.analysis_strided(x, filt, self.hop_size)
        else:
            # conv2d stride 1 + average_pooling2d(pool, stride pool) (adapt.py:118-120); trainable through the box-filtered
            # signal (layers._AnalysisAvgFn); no arg-max: back() places the atoms at the fixed positions tp*pool
            y, am = L.analysis_avg(x, filt, self.max_pool_value), None
        return y, am

    # adapt.py:162-196 (pretraining separator)
    def separator(self, y, B):
        S = self.S
        Tp, N = y.shape[1], y.shape[2]
        input_mix = y[:B].reshape(B, 1, Tp, N)
        input_non_mix = y[B:].reshape(B, S, Tp, N)
        if self.separation == "mask":
            out = input_mix * (input_non_mix / input_mix)
        else:
            out = input_mix - (input_non_mix.sum(1, keepdim=True) - input_non_mix)
        return out.reshape(B * S, Tp, N)

    # adapt.py:205-252
    def back(self, sep_out, argmax, B, Lw):
        filt2 = self.conv_filter("back")
        if self.with_max_pool:
            out = L.synthesis(sep_out.contiguous(), argmax[:B].contiguous(), filt2, B, self.S, Lw, self.max_pool_value,
                              self.hop_size)
        elif not self.with_average_pool:
            # conv2d_transpose with strides [1,1,hop,1] (adapt.py:236-243) = overlap-add of atoms at the fixed positions
            _, c = L.strided_positions(1, Lw, self.window, self.N, self.hop_size, sep_out.device)
            out = L.synthesis(sep_out.contiguous(), argmax[:B].contiguous(), filt2, B, self.S, Lw, c + 1, self.hop_size)
        else:
            # UpSampling2D((1, pool)) + conv2d_transpose stride 1 (adapt.py:224-243) = sparse overlap-add of the box-filtered bank
            out = L.synthesis_avg(sep_out.contiguous(), filt2, B, self.S, Lw, self.max_pool_value)
        return out.reshape(B, self.S, Lw)

    # network.py:196-221 (with_perm=False): SDR improvement metric and the 'sdr' loss ratio per (b,s)
    def sdr_improvement(self, x_mix, s_target, s_approx):
        B, S, Lw = s_target.shape
        st = L.wave_stats(s_target.reshape(B * S, Lw), s_approx.reshape(B * S, Lw))
        tn, an, ta = st[:, 0], st[:, 1], st[:, 2]
        with torch.no_grad():
            mixr = x_mix.unsqueeze(1).expand(B, S, Lw).reshape(B * S, Lw).contiguous()
            sm = ops.wave_stats(s_target.reshape(B * S, Lw).contiguous(), mixr)
            separated = 10.0 * torch.log(1.0 / ((tn * an) / (ta * ta) - 1.0)) / math.log(10.0)
            non_sep = 10.0 * torch.log(1.0 / ((sm[:, 0] * sm[:, 1]) / (sm[:, 2] * sm[:, 2]) - 1.0)) / math.log(10.0)
            val = (separated - non_sep).reshape(B, S).mean(-1).mean(-1)
        loss = (tn * an) / (ta * ta + 1e-12)
        return val, loss.reshape(B, S), st

    # adapt.py:141-160
    def overlap(self, y, B):
        S = self.S
        nm = y[B:].reshape(B, S, -1).abs()
        vals = []
        for a, b in itertools.combinations(range(S), 2):
            pa, pb = nm[:, a], nm[:, b]
            vals.append((1.0 - (pa - pb).abs() / (torch.maximum(pa, pb) + 1e-8)).mean(-1))
        return torch.stack(vals, 1).mean(-1).mean(-1)

    # adapt.py:307-338, 374-385 (pretraining branch)
    def cost(self, x_mix, x_non_mix):
        """Pre-training cost of the autoencoder: l2 / sdr (/ both) + beta*KL sparsity + lambda^2*reg +
        overlap_coef*overlap + non_negativity^2*neg.  Every node is a library kernel: analysis, the fused front-output
        terms (p_hat / KL, overlap, separator, non-negativity: amss_adapt_terms_*), synthesis, waveform statistics, and
        their gradients; what remains on the host side is scalar arithmetic on a handful of device scalars.
        Returns (cost, aux)."""
        B, S, Lw = x_non_mix.shape
        y, am = self.front(x_mix, x_non_mix)
        filt, filt2 = self.conv_filter("front"), self.conv_filter("back")
        sep, p_hat, terms = L.adapt_terms(y, B, S, self.p, 0 if self.separation == "mask" else 1)
        sparse, overlapping, nonneg = terms[0], terms[1], terms[2]
        back = self.back(sep, am, B, Lw)
        # l2 (adapt.py:323-325), sdr (:327-330), the choice of loss (:332-337), beta * KL, lambda^2 * reg (lambda applied
        # twice, :312, :380), overlap_coef * overlap, nn^2 * neg (applied twice, :316, :384) and the SDR-improvement metric
        # (network.py:196-221) from the waveform statistics: one kernel (amss_adapt_cost_fwd)
        tgt = x_non_mix.reshape(B * S, Lw)
        st = L.wave_stats(tgt, back.reshape(B * S, Lw))
        with torch.no_grad():
            sm = ops.wave_stats_rows(tgt.contiguous(), x_mix.contiguous(), S)
        cost, aux3 = L.adapt_cost(st, terms, filt, filt2, sm, B, S, self.loss, self.beta, self.l, self.overlap_coef,
                                  float(self.non_negativity or 0.0))
        l2, sdr, val = aux3[0], aux3[1], aux3[2]
        return cost, {"y": y, "argmax": am, "back": back, "l2": l2, "sdr": sdr, "sdr_improvement": val,
                      "sparse_constraint": sparse, "overlapping": overlapping, "p_hat": p_hat}

    # adapt.py:339-372 + network.py:196-221 (with_perm=True): the non-pretraining branch of Adapt.cost
    def cost_separation(self, x_mix, x_non_mix, back, front_y=None):
        """Adapt.cost with pretraining=False: `back` [B,S,L] is the synthesis of the separator's output.
          l2  = mean_B min_perm sum_S mean_L (x_s - back_perm(s))^2                              (adapt.py:353-356)
          sdr = mean_B sum_S min_{b'} tn[b,s] an[b',s] / (<x[b,s], back[b',s]>^2 + 1e-12)          (:358-362)
        The sdr term reproduces the reference as written: it hands the UN-permuted `back` [B,S,L] to sdr_improvement next to
        targets shaped [B,1,S,L], so broadcasting pairs target b with estimate b' of every OTHER mixture and the
        `reduce_min(sdr, 1)` that was meant to run over permutations runs over b' (for B = 1 it is the plain ratio).
        loss = l2 | sdr | 1e-3 * l2 + sdr (:364-369), then the same regularisers as the pretraining branch (:374-385; the KL
        and overlap terms need the front output: pass front_y).  The pairwise <x[b,s], back[b',s]> come from the library
        GEMM, everything else is a [B,B,S] table.  Returns (cost, aux)."""
        B, S, Lw = x_non_mix.shape
        l2 = 2.0 / Lw * L.pit_wave_l2(x_non_mix, back, reduce="sum")                       # 0.5*sum -> mean over L, sum over S
        tn = (x_non_mix ** 2).sum(-1)                                                      # [B,S]   (data, no gradient)
        st = L.wave_stats(back.reshape(B * S, Lw), back.reshape(B * S, Lw))                # <a,a> with autograd
        an = st[:, 1].reshape(B, S)
        cross = torch.stack([L.pair_dots(x_non_mix[:, s].contiguous(), back[:, s].contiguous()) for s in range(S)], 2)  # [B,B',S]
        ratio = tn.unsqueeze(1) * an.unsqueeze(0) / (cross ** 2 + 1e-12)                   # [B,B',S]
        sdr = ratio.min(1).values.sum(-1).mean(-1)
        with torch.no_grad():                                                              # SDR improvement metric (with_perm)
            mixn = (x_mix ** 2).sum(-1)                                                    # [B'] (mix is tiled to [1,B',S,L])
            cm = torch.stack([L.pair_dots(x_non_mix[:, s].contiguous(), x_mix.contiguous()) for s in range(S)], 2)
            separated = 10.0 * torch.log(1.0 / ((tn.unsqueeze(1) * an.unsqueeze(0)) / cross ** 2 - 1.0)) / math.log(10.0)
            non_sep = 10.0 * torch.log(1.0 / ((tn.unsqueeze(1) * mixn.view(1, B, 1)) / cm ** 2 - 1.0)) / math.log(10.0)
            val = (separated - non_sep).mean(-1).mean(0).max(-1).values
        cost = l2 if self.loss == "l2" else (sdr if self.loss == "sdr" else 1e-3 * l2 + sdr)
        filt, filt2 = self.conv_filter("front"), self.conv_filter("back")
        if front_y is not None:
            _, _, terms = L.adapt_terms(front_y, B, S, self.p, 1, want_sep=False)
            if self.beta != 0.0:
                cost = cost + self.beta * terms[0]
            if self.overlap_coef != 0.0:
                cost = cost + self.overlap_coef * terms[1]
            if self.non_negativity:
                cost = cost + self.non_negativity * (self.non_negativity * terms[2])
        if self.l != 0.0:
            cost = cost + self.l * (self.l * (0.5 * (filt2 ** 2).sum() + 0.5 * (filt ** 2).sum()))
        return cost, {"l2": l2, "sdr": sdr, "sdr_improvement": val}

    # adapt.py:404-431
    def cost_finetuning(self, x_non_mix, back):
        """PIT waveform loss of the end-to-end fine-tuning recipes: back [B,S,L] = self.back(sepNet output)."""
        return L.pit_wave_l2(x_non_mix, back)

    def connect_front(self, separator_class, **extra):
        """adapt.py:440-441 -- plug a Separator subclass on the front output (plugged=True)."""
        args = dict(self.args)
        args.update(extra)
        self.sepNet = separator_class(plugged=True, store=self.store, F=self.N, **args)
        return self.sepNet


# =================================================================================================
# Separator (STFT or plugged on the adaptive front)                          models/network.py:313-723
# =================================================================================================
class Separator(Network):
    mask_a, mask_b = 1.0, 0.0

    def __init__(self, plugged=False, store=None, F=None, **kwargs):
        kwargs.pop("mask_a", None), kwargs.pop("mask_b", None)
        if not plugged:
            kwargs.setdefault("window_size", 512)
            kwargs.setdefault("hop_size", 256)
        super().__init__(store, **kwargs)
        a = self.args
        self.plugged = plugged
        self.window_size, self.hop_size = a["window_size"], a["hop_size"]
        self.F = F if plugged else self.window_size // 2 + 1
        self.layer_size, self.embedding_size = a["layer_size"], a["embedding_size"]
        self.normalize = a["no_normalize"]          # store_false flag: True = normalise (network.py:322)
        self.nb_layers = a["nb_layers"]
        self.num_speakers = a["tot_speakers"]
        self.beta, self.threshold = a["beta_kmeans"], a["threshold"]
        self.with_silence, self.nb_tries, self.nb_steps = a["with_silence"], a["nb_tries"], a["nb_steps"]
        self.abs_input = a["abs_input"]
        # input / label options (models/network.py:381-396, 409-443, 504-521): library kernels amss_separator_input_prep /
        # amss_label_weights; argparse hands over the STRING 'None' for the unset choices
        self.normalize_input, self.pre_func = a["normalize_separator"], a["pre_func"]
        self.silent_threshold = a["silence_mask_db"]
        self.function_mask, self.loss_with_silence = a["function_mask"], a["silence_loss"]
        self.threshold_silence_loss = a["threshold_silence_loss"]
        for flag, table in (("normalize_separator", ops.NORMALIZE), ("pre_func", ops.PRE_FUNC), ("function_mask", ops.FUNCTION_MASK)):
            if a[flag] not in table:
                raise ValueError(f"--{flag} {a[flag]!r}: the reference knows {sorted(k for k in table if k)}")
        if a["add_dilated"] or a["sampling"] is not None:
            raise NotImplementedError("add_dilated / negative sampling are outside the hot path (SURVEY section 2, rows 2 and 4)")
        if not plugged and (self.function_mask not in ("None", None) or self.loss_with_silence):
            raise ValueError("--function_mask / --silence_loss act on the plugged separator only (models/network.py:381-396)")
        self._build_prediction()

    @property
    def weighted_labels(self):
        return self.plugged and (self.function_mask not in ("None", None) or bool(self.loss_with_silence))

    def _prep(self, X, plugged):
        """init_separator (models/network.py:409-443): plugged: abs_input -> normalisation; STFT: pre_func -> normalisation
        -> silent-dB mask.  One library kernel; no-op (and no launch) with the default flags."""
        if plugged:
            on = self.abs_input or self.normalize_input not in ("None", None)
            if not on:
                return X
            if X.requires_grad:
                raise AmssError("separator input options are forward-only: the front end must be frozen")
            return ops.separator_input_prep(X.contiguous(), abs_input=self.abs_input, normalize=self.normalize_input)
        on = self.pre_func not in ("None", None) or self.normalize_input not in ("None", None) or self.silent_threshold > 0
        if not on:
            return X
        return ops.separator_input_prep(X.contiguous(), pre_func=self.pre_func, normalize=self.normalize_input,
                                        silence_db=float(self.silent_threshold))

    # DPCL.prediction / L41Model.prediction trunk (dpcl.py:19-39, L41.py:21-45)
    def _build_prediction(self):
        st, prec = self.store, self.precision
        in_dim = self.F
        self.layers = []
        for i in range(self.nb_layers):
            self.layers.append(L.BLSTM(self.layer_size, name=f"BLSTM_{i}", drop_val=self.args["recurrent_dropout"],
                                       store=st, scope="prediction", in_dim=in_dim, precision=prec))
            in_dim = 2 * (self.layer_size // 2)
        self.layers.append(L.Conv1D([1, in_dim, self.embedding_size * self.F], store=st, scope="prediction",
                                    precision=prec, reference_scale=self.args["reference_init"]))

    # network.py:480-502
    def preprocessing(self, x_mix, x_non_mix, want_mag_non_mix=False):
        spec, X = ops.stft(x_mix.contiguous(), self.window_size, self.hop_size)
        labels, mag_nm = ops.stft_labels(x_non_mix.contiguous(), self.window_size, self.hop_size, want_mag_non_mix)
        # X_input (the magnitudes the masks multiply, network.py:498) stays raw; X (the network input) takes the options
        return {"stfts": spec, "X": self._prep(X, False), "X_input": X, "labels": labels, "X_non_mix": mag_nm}

    # network.py:357-400 (plugged branch, default flags)
    def plugged_inputs(self, front_y, B):
        X = front_y[:B]
        labels = ops.plugged_labels(front_y.contiguous(), B, self.S)
        weights = None
        if self.weighted_labels:                             # y * f(|X| / max) and / or y * silence mask (network.py:381-396)
            weights = ops.label_weights(X.detach().contiguous(), self.function_mask,
                                        self.threshold_silence_loss if self.loss_with_silence else 0.0)
        return {"X": self._prep(X, True), "labels": labels, "X_raw": X, "weights": weights}

    def prediction(self, X):
        """X [B,T,F] -> embeddings [B,T,F,E] (L2-normalised over E unless --no_normalize)."""
        B, Tt, Fb = X.shape
        E, head = self.embedding_size, self.layers[-1]
        if (self.normalize and self.precision != L.AMSS_PREC_FP32 and isinstance(head, L.Conv1D)
                and E % 8 == 0 and E <= 48):
            # tensor-core path: Reshape + Normalize run inside the head GEMM's epilogue
            V = head.f_prop_normalized(L.f_props(self.layers[:-1], X), E)
            return L._carry(V, V.view(B, Tt, Fb, E))
        z = L.f_props(self.layers, X)
        z = L.Reshape([B, Tt, Fb, self.embedding_size]).f_prop(z)
        return L.Normalize(3).f_prop(z) if self.normalize else z

    # network.py:554-582
    def separate(self, V, X_input, init_idx=None, rng=None):
        """V [B,T,F,E], X_input [B,T,F] -> (separated [B*S,T,F], labels int32 [B,TF] | soft [B,TF,S])."""
        B, Tt, Fb, E = V.shape
        emb = V.detach().reshape(B, Tt * Fb, E).contiguous()
        if rng is None:               # one stream per separator (fresh draws every call, as the reference's py_func), per rank
            if not hasattr(self, "_kmeans_rng"):
                import os
                self._kmeans_rng = np.random.RandomState(self.args["seed"] + int(os.environ.get("RANK", "0")))
            rng = self._kmeans_rng
        km = KMeans(nb_clusters=self.S, nb_tries=self.nb_tries, nb_iterations=self.nb_steps, beta=self.beta,
                    latent_space_tensor=X_input.abs().reshape(B, Tt * Fb) if self.with_silence else None,
                    threshold=self.threshold, assign_at_end=self.args["end_assign"], rng=rng)
        _, lab = km.fit(emb, init_idx)
        self.kmeans = km
        Xf = X_input.reshape(B, Tt * Fb).contiguous()
        sep = ops.apply_masks(Xf, self.S, labels=lab) if self.beta is None else ops.apply_masks(Xf, self.S, soft=lab)
        return sep.view(B * self.S, Tt, Fb), lab

    # network.py:456-462, 629-636: the enhance BLSTM stack on [separated || X] (config 3)
    def add_enhance_layer(self):
        a, st, prec = self.args, self.store, self.precision
        in_dim = 2 * self.F
        self.enhance_layers = []
        for i in range(a["nb_layers_enhance"]):
            self.enhance_layers.append(L.BLSTM(a["layer_size_enhance"], name=f"BLSTM_{i}",
                                               drop_val=a["recurrent_dropout_enhance"], store=st, scope="enhance",
                                               in_dim=in_dim, precision=prec))
            in_dim = 2 * (a["layer_size_enhance"] // 2)
        self.enhance_layers.append(L.Conv1D([1, in_dim, self.F], store=st, scope="enhance", precision=prec))
        return self

    # network.py:610-639: [separated || X_input] -> enhance BLSTM stack -> Conv1D: the logits [B,S,TF]
    def enhance_logits(self, separated, X_input):
        B, Tt, Fb = X_input.shape
        S = self.S
        sep4 = separated.reshape(B, S, Tt, Fb)
        z = torch.cat([sep4, X_input.unsqueeze(1).expand(B, S, Tt, Fb)], 3).reshape(B * S, Tt, 2 * Fb)
        if self.args["normalize_enhance"]:
            mean = z.mean((1, 2), keepdim=True)
            var = z.var((1, 2), unbiased=False, keepdim=True)
            z = (z - mean) / torch.sqrt(var)
        yv = L.f_props(self.enhance_layers, z.contiguous())                       # [B*S,T,F]
        return yv.reshape(B, S, Tt * Fb)

    # network.py:640-693 fused: nonlinearity over the sources, * X_input, PIT-L2 against the sources' magnitudes
    def enhance_cost_fused(self, logits, X_input, X_non_mix):
        B, S, TF = logits.shape
        return L.enhance_cost_fused(logits, X_input.reshape(B, TF), X_non_mix.reshape(B, TF, S), self.args["nonlinearity"])

    # network.py:610-660
    def enhance(self, separated, X_input):
        """separated [B*S,T,F] (k-means masks applied), X_input [B,T,F] -> (enhanced [B,S,TF], cost_in [B,TF,S])."""
        B, Tt, Fb = X_input.shape
        S = self.S
        yv = self.enhance_logits(separated, X_input).transpose(1, 2)             # [B,TF,S]
        nl = self.args["nonlinearity"]
        if nl == "softmax":
            yv = torch.softmax(yv, -1)
        elif nl == "tanh":
            yv = torch.tanh(yv)
        self.enhance_masks = yv                                                  # [B,TF,S]: postprocessing_masks() input
        cost_in = yv * X_input.reshape(B, -1, 1)
        return cost_in.transpose(1, 2), cost_in

    # network.py:662-693: PIT L2 over the S! permutations, min over perms, mean over the batch
    def enhance_cost(self, cost_in, X_non_mix):
        B, TF, S = cost_in.shape
        est = cost_in.transpose(1, 2)
        tgt = X_non_mix.reshape(B, TF, S).transpose(1, 2)
        costs = [((tgt - est[:, list(perm)]) ** 2).sum(-1).sum(-1) for perm in itertools.permutations(range(S))]
        return torch.stack(costs, 1).min(1).values.mean()

    # network.py:697-723
    def cost_finetuning(self, x_non_mix, postprocessed):
        """PIT waveform loss on the separated waveforms [B,S,L] (same definition as Adapt.cost_finetuning)."""
        return L.pit_wave_l2(x_non_mix, postprocessed)

    def postprocessing_masks(self, stfts, masks):
        """Differentiable postprocessing for the fine-tuning recipes: masks [B,TF,S] (e.g. the enhance layer's output
        ratio enhanced/|X|) -> waveforms [B,S,L'] with autograd to the masks."""
        return L.istft_masked(stfts, masks, self.S, self.window_size, self.hop_size)

    # network.py:584-607
    def postprocessing(self, stfts, labels_or_masks):
        if self.beta is None:
            return ops.istft_masked(stfts, self.S, self.window_size, self.hop_size, labels=labels_or_masks)
        return ops.istft_masked(stfts, self.S, self.window_size, self.hop_size, masks=labels_or_masks)


class DPCL(Separator):
    """models/dpcl.py -- affinity loss with un-squared Frobenius norms, one-hot labels 1/0."""
    mask_a, mask_b = 1.0, 0.0

    def cost(self, V, labels, I=None, weights=None):
        B, Tt, Fb, E = V.shape
        if weights is not None and self.loss_with_silence:
            # the silence mask zeroes rows of Y, and the reference's D = 1/sqrt(Y Y^T 1) (dpcl.py:58-62) is then 1/sqrt(0):
            # its DPCL cost is inf / NaN from the first step.  Refused loudly instead of reproducing the NaN
            raise ValueError("--silence_loss with the DPCL cost divides by zero in the reference (models/dpcl.py:58-62); "
                             "use it with L41, or --function_mask alone")
        # weights: Y = one_hot * f(|X| / max) (--function_mask, network.py:381-389), fp32 kernels
        return L.dpcl_loss(V.reshape(B, Tt * Fb, E), labels.reshape(B, Tt * Fb), self.S,
                           prenorm=getattr(V, "_amss_prenorm", None), precision=self.precision,
                           head=getattr(V, "_amss_head", None), weights=weights)


class L41Model(Separator):
    """models/L41.py -- speaker-vector table [tot_speakers, E], sigmoid-dot loss, labels +1/-1."""
    mask_a, mask_b = 1.0, -1.0

    def __init__(self, plugged=False, store=None, F=None, **kwargs):
        super().__init__(plugged, store, F, **kwargs)
        E = self.embedding_size
        g = self.store.gen
        v = torch.randn(self.num_speakers, E, generator=g, dtype=torch.float64).clamp_(-2.0, 2.0) * math.sqrt(2.0 / E)
        self.store.register("speaker_centroids", v.float())          # L41.py:16-18

    def cost(self, V, labels, I, weights=None):
        B, Tt, Fb, E = V.shape
        sv = self.store["speaker_centroids"]
        if self.normalize:
            sv = L.l2_normalize(sv, E)                               # L41.py:60-61
        spk = sv[I.long()].contiguous()                              # [B,S,E] gather (L41.py:62)
        return L.l41_loss(V.reshape(B, Tt * Fb, E), labels.reshape(B, Tt * Fb), spk, weights)


# =================================================================================================
# KMeans                                                                       models/Kmeans_2.py
# =================================================================================================
_IDX_RINGS = {}      # pinned staging rings of the k-means initial rows, keyed by shape (KMeans.fit)


class KMeans:
    """KMeans(nb_clusters, centroids_init, nb_tries, nb_iterations, input_tensor, normalize_input,
    latent_space_tensor, beta, threshold, assign_at_end) -- same argument names as the reference
    (Kmeans_2.py:14-17).  The random initial rows are drawn on the host exactly as the reference's
    py_func does (np.random.choice without replacement per row, :61-65) unless init_idx is given."""

    def __init__(self, nb_clusters, centroids_init=None, nb_tries=10, nb_iterations=10, input_tensor=None,
                 normalize_input=True, latent_space_tensor=None, beta=None, threshold=2.5, assign_at_end=True,
                 rng=None):
        if centroids_init is not None:
            raise NotImplementedError("centroids_init is dead code in the reference (Kmeans_2.py:72-74)")
        self.nb_clusters, self.nb_tries, self.nb_iterations = nb_clusters, nb_tries, nb_iterations
        self.normalize_input, self.latent, self.beta = normalize_input, latent_space_tensor, beta
        self.threshold, self.assign_at_end = threshold, assign_at_end
        self.rng = rng if rng is not None else np.random.RandomState(42)
        self.input_tensor = input_tensor

    def random_init(self, rows, Lp):
        """`nb_clusters` distinct rows of X per (mixture, try), uniformly at random -- the distribution of the reference's
        py_func (np.random.choice(range(l), size=K, replace=False) per row, Kmeans_2.py:61-65), drawn for all rows at once:
        np.random.choice without replacement permutes the whole population per call (O(L) each; 0.5 s of host time per
        64-mixture batch at TF = 64000), so rows are drawn with replacement and the few rows that repeat an index are
        redrawn."""
        K = self.nb_clusters
        idx = self.rng.randint(0, Lp, size=(rows, K))
        while True:
            srt = np.sort(idx, axis=1)
            bad = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
            if not bad.any():
                return idx.astype(np.int32)
            idx[bad] = self.rng.randint(0, Lp, size=(int(bad.sum()), K))

    def fit(self, X, init_idx=None):
        """X [B,L,E] (CUDA) -> (centroids [B,K,E], labels int32 [B,L] or soft [B,L,K])."""
        B, Lp, E = X.shape
        if init_idx is None:
            init_idx = self.random_init(B * self.nb_tries, Lp)
        if torch.is_tensor(init_idx) and init_idx.is_cuda:
            idx = init_idx.to(torch.int32).contiguous()
        else:
            # pinned staging + asynchronous copy: a pageable .to(device) blocks the host until the GPU has drained everything
            # queued before it (the whole trunk of this step), which serialises a streaming loop
            host = torch.as_tensor(np.asarray(init_idx), dtype=torch.int32).contiguous()
            # (the ring is shared by every KMeans of the process: Separator.separate builds a new object per call, and pinning
            # costs ~0.1 ms per buffer)
            ring = _IDX_RINGS.get(tuple(host.shape))
            if ring is None:
                ring = _IDX_RINGS[tuple(host.shape)] = {"bufs": [[torch.empty(host.shape, dtype=torch.int32).pin_memory(), None]
                                                                 for _ in range(4)], "k": 0}
            buf = ring["bufs"][ring["k"]]
            ring["k"] = (ring["k"] + 1) % len(ring["bufs"])
            if buf[1] is not None:
                buf[1].synchronize()                     # the copy that last used this staging buffer has run
            buf[0].copy_(host)
            idx = buf[0].to(X.device, non_blocking=True)
            buf[1] = torch.cuda.Event()
            buf[1].record(torch.cuda.current_stream())
        ns = None
        if self.latent is not None:
            ns = ops.kmeans_silence_mask(self.latent.reshape(B, Lp).contiguous(), self.threshold)
        cent, lab, inertia, best = ops.kmeans_fit(X.contiguous(), idx, self.nb_clusters, self.nb_tries,
                                                  self.nb_iterations, self.beta, ns, self.normalize_input,
                                                  self.assign_at_end)
        self.inertia, self.best_try = inertia, best
        return cent, lab

    @property
    def network(self):
        return self.fit(self.input_tensor)
