"""Training / inference recipes of the hot path (reference utils/trainer.py): the flag parser
(MyArgs, :10-176), the trainer subclasses whose build() picks the model and freeze recipe
(:392-658) and the train loop (:264-390), re-expressed as one process per GPU with a single NCCL
all-reduce of the flat gradient buffer per step.  The data layer is out of scope:
any iterator of (mix [B,L], non_mix [B,S,L], ind [B,S]) host arrays can be fed."""
import argparse
import os
import time

import numpy as np
import torch

from . import _lib
from . import dp, ops
from .models import Adapt, DPCL, L41Model, DEFAULTS  # noqa: F401


class MyArgs:
    """Same flag names and defaults as the reference's MyArgs (utils/trainer.py:10-176)."""

    def __init__(self):
        p = argparse.ArgumentParser(description="Argument Parser")
        p.add_argument("--dataset_normalize", action="store_true")
        p.add_argument("--chunk_size", type=int, default=20480)
        p.add_argument("--nb_speakers", type=int, default=2)
        p.add_argument("--validation_step", type=int, default=1000)
        p.add_argument("--epochs", type=int, default=10)
        p.add_argument("--batch_size", type=int, default=64)
        p.add_argument("--learning_rate", type=float, default=0.1)
        p.add_argument("--optimizer", choices=["Adam", "SGD", "RMSProp"], default="Adam")
        p.add_argument("--decay_epoch", type=int, default=50)
        p.add_argument("--gradient_norm_clip", type=float, default=0.0)
        p.add_argument("--precision", choices=["fp32", "bf16"], default="fp32")
        self.parser = p

    def add_stft_args(self):
        self.parser.add_argument("--window_size", type=int, default=512)
        self.parser.add_argument("--hop_size", type=int, default=256)

    def add_separator_args(self):
        p = self.parser
        p.add_argument("--normalize_separator", choices=["None", "01", "meanstd"], default="None")
        p.add_argument("--abs_input", action="store_true")
        p.add_argument("--pre_func", choices=["None", "sqrt", "log"], default="None")
        p.add_argument("--silence_mask_db", type=int, default=0)
        p.add_argument("--nb_layers", type=int, default=3)
        p.add_argument("--layer_size", type=int, default=600)
        p.add_argument("--embedding_size", type=int, default=40)
        p.add_argument("--no_normalize", action="store_false")
        p.add_argument("--recurrent_dropout", type=float, default=0.0)
        p.add_argument("--nb_tries", type=int, default=10)
        p.add_argument("--nb_steps", type=int, default=10)
        p.add_argument("--beta_kmeans", type=float, default=None)
        p.add_argument("--threshold", type=float, default=2.0)
        p.add_argument("--with_silence", action="store_true")
        p.add_argument("--end_assign", action="store_true")
        p.add_argument("--silence_loss", action="store_true")
        p.add_argument("--threshold_silence_loss", type=float, default=2.0)
        p.add_argument("--function_mask", choices=["None", "linear", "sqrt", "square"], default="None")
        p.add_argument("--sampling", type=int, default=None)
        p.add_argument("--ns_rate", type=float, default=0.1)
        p.add_argument("--ns_method", choices=["random", "k-nearest"], default="random")
        p.add_argument("--add_dilated", action="store_true")

    def add_adapt_args(self):
        p = self.parser
        p.add_argument("--window_size", type=int, default=1024)
        p.add_argument("--filters", type=int, default=512)
        p.add_argument("--max_pool", type=int, default=512)
        p.add_argument("--with_max_pool", action="store_true")
        p.add_argument("--with_average_pool", action="store_true")
        p.add_argument("--hop_size", type=int, default=256)
        p.add_argument("--regularization", type=float, default=1e-4)
        p.add_argument("--beta", type=float, default=1e-2)
        p.add_argument("--sparsity", type=float, default=0.01)
        p.add_argument("--overlap_coef", type=float, default=0.001)
        p.add_argument("--overlap_value", type=float, default=0.1)
        p.add_argument("--non_negativity", type=float, default=0.0)
        p.add_argument("--loss", choices=["l2", "sdr", "l2+sdr", "sdr+l2"], default="sdr")
        p.add_argument("--separation", choices=["perfect", "mask"], default="perfect")

    def get_args(self, argv=None):
        return vars(self.parser.parse_args(argv))


class AMSGradOptimizer:
    """Network.optimize (models/network.py:167-194): 'Adam' means AMSGrad(lr, beta1=0.9, beta2=0.99,
    epsilon=1e-3) with a constant learning rate (utils/ops.py:639-704) and an optional global-norm
    clip.  One fused kernel over the flat parameter range."""

    def __init__(self, store, lr, beta1=0.9, beta2=0.99, eps=1e-3, clip=0.0):
        self.store, self.lr, self.b1, self.b2, self.eps, self.clip = store, lr, beta1, beta2, eps, clip
        n = store.n_trainable
        self.m = torch.zeros(n, dtype=torch.float32, device=store.device)
        self.v = torch.zeros_like(self.m)
        self.vhat = torch.zeros_like(self.m)
        self.t = 0

    def step(self, grad_scale=1.0):
        st = self.store
        self.t += 1
        lr_t = ops.amsgrad_lr_t(self.lr, self.b1, self.b2, self.t)
        fac = ops.global_norm_clip_factor(st.grad_flat, self.clip) if self.clip else None
        ops.amsgrad_step(st.flat[:st.n_trainable], st.grad_flat, self.m, self.v, self.vhat, lr_t, self.b1, self.b2,
                         self.eps, grad_scale, fac)


class DevicePrefetcher:
    """Host -> device input staging one step ahead (the role of the reference's tf.data prefetch(1),
    data/dataset.py:506-514): pinned host batches are copied on a side stream into double-buffered device
    inputs while the previous step computes; stream events (no host sync) order copy and compute."""

    def __init__(self, example_batch):
        self.stream = torch.cuda.Stream()
        self.bufs = [[torch.empty(tuple(a.shape), dtype=a.dtype, device="cuda") for a in example_batch] for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [None, None]
        self.k = 0

    @staticmethod
    def pin(batch):
        return tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in batch)

    def submit(self, pinned_batch):
        """Start the H2D copy of a pinned batch; returns the slot to pass to get()."""
        k = self.k
        self.k ^= 1
        if self.done[k] is not None:
            self.stream.wait_event(self.done[k])          # the step that last read this slot has finished
        with torch.cuda.stream(self.stream):
            for d, h in zip(self.bufs[k], pinned_batch):
                d.copy_(h, non_blocking=True)
            self.ready[k].record(self.stream)
        return k

    def get(self, k):
        torch.cuda.current_stream().wait_event(self.ready[k])
        return self.bufs[k]

    def release(self, k):
        """Call after launching the step that consumes slot k."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.done[k] = ev


class Trainer:
    """Common loop (utils/trainer.py:264-390).  Subclasses implement build() and loss(batch)."""

    def __init__(self, **kwargs):
        self.args = dict(kwargs)
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.distributed = self.world > 1 and torch.distributed.is_available() and torch.distributed.is_initialized()
        self.build()
        self.store.finalize()
        self.post_build()
        self.optimizer = AMSGradOptimizer(self.store, self.args.get("learning_rate", DEFAULTS["learning_rate"]),
                                          clip=self.args.get("gradient_norm_clip", 0.0))
        self._pinned = None

    # -- to override ---------------------------------------------------------------------------
    def build(self):
        raise NotImplementedError

    def post_build(self):
        pass

    def loss(self, x_mix, x_non_mix, ind):
        raise NotImplementedError

    # -- host -> device staging from pinned memory ----------------------------------------------
    def to_device(self, batch):
        mix, non_mix, ind = batch
        if self._pinned is None or self._pinned[0].shape != mix.shape:
            self._pinned = tuple(torch.empty(a.shape, dtype=torch.from_numpy(np.asarray(a)).dtype).pin_memory()
                                 for a in (mix, non_mix, ind))
        out = []
        for buf, a in zip(self._pinned, (mix, non_mix, ind)):
            buf.copy_(torch.from_numpy(np.ascontiguousarray(a)))
            out.append(buf.to("cuda", non_blocking=True))
        return out

    # -- optional CUDA-graph replay of forward + backward ---------------------------------------
    def enable_cuda_graph(self, eager_steps=2):
        """Capture zero-grad + loss + backward of one step into a CUDA graph (after `eager_steps` ordinary steps, which
        also warm every lazily initialised kernel attribute) and replay it from then on; the gradient all-reduce and the
        optimizer stay outside the graph (their arguments change per step).  Each train_step() call is still exactly
        one optimisation step on the batch it is given (inputs are copied into the graph's static buffers).  The loss
        must not depend on host-side state that changes between steps (k-means initial rows, Python control flow on
        data): use it for the Front / STFT separator trainers."""
        self._cg = {"eager": int(eager_steps), "calls": 0, "graph": None, "inputs": None, "cost": None}

    def graph_kernel_launches(self):
        """Library kernels executed through graph replays so far (they do not pass through the host-side launch counter)."""
        cg = getattr(self, "_cg", None)
        return cg["kernels"] * cg["replays"] if cg and cg["graph"] is not None else 0

    def _capture_step(self, x_mix, x_non_mix, ind):
        cg = self._cg
        cg["inputs"] = [t.clone() for t in (x_mix, x_non_mix, ind)]
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            self.store.grad_flat.zero_()
            cost = self.loss(*cg["inputs"])
            cost.backward()
        cg["graph"], cg["cost"] = graph, cost.detach()
        cg["kernels"] = _lib.launch_count() - n0          # library kernels recorded in the graph = run by every replay
        cg["replays"] = 0

    # -- one optimisation step on device tensors ------------------------------------------------
    def train_step(self, x_mix, x_non_mix, ind):
        cg = getattr(self, "_cg", None)
        if cg is not None:
            if cg["graph"] is None and cg["calls"] >= cg["eager"]:
                self._capture_step(x_mix, x_non_mix, ind)
            cg["calls"] += 1
            if cg["graph"] is not None:
                for dst, src in zip(cg["inputs"], (x_mix, x_non_mix, ind)):
                    dst.copy_(src, non_blocking=True)
                cg["graph"].replay()
                cg["replays"] += 1
                scale = dp.allreduce_sum_(self.store.grad_flat)
                self.optimizer.step(scale)
                return cg["cost"].clone()
        self.store.grad_flat.zero_()
        cost = self.loss(x_mix, x_non_mix, ind)
        cost.backward()
        scale = dp.allreduce_sum_(self.store.grad_flat)            # the single collective of the path
        self.optimizer.step(scale)
        return cost.detach()

    def train(self, data, steps, log_every=0):
        """data: iterator of host batches (numpy or pinned tensors).  Inputs are staged one step ahead
        (DevicePrefetcher) and each step's cost is read back while the next step runs.
        Returns the list of per-step costs (floats)."""
        costs = []
        t0 = time.time()
        as_pinned = lambda b: b if torch.is_tensor(b[0]) and b[0].is_pinned() else DevicePrefetcher.pin(b)  # noqa: E731
        first = as_pinned(next(data))
        pf = DevicePrefetcher(first)
        slot, pending = pf.submit(first), None
        for step in range(steps):
            c = self.train_step(*pf.get(slot))
            pf.release(slot)
            if step + 1 < steps:
                slot = pf.submit(as_pinned(next(data)))
            if pending is not None:
                costs.append(float(pending))
            pending = c
            if log_every and (step + 1) % log_every == 0 and self.rank == 0 and costs:
                print(f"step {step + 1}/{steps} loss={costs[-1]:.6f} {(time.time() - t0) / (step + 1):.3f} s/step")
        costs.append(float(pending))
        return costs


class STFT_Separator_Trainer(Trainer):
    """utils/trainer.py:468-486 -- STFT + DPCL / L41 (BASELINE config 1 / 3 trunk)."""

    def __init__(self, separator, name="STFT_Separator", **kwargs):
        self.separator_class, self.name = separator, name
        super().__init__(**kwargs)

    def build(self):
        args = {k: v for k, v in self.args.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        self.model = self.separator_class(plugged=False, **args)
        self.store = self.model.store

    def loss(self, x_mix, x_non_mix, ind):
        pre = self.model.preprocessing(x_mix, x_non_mix)
        V = self.model.prediction(pre["X"])
        return self.model.cost(V, pre["labels"], ind)


class Front_Separator_Trainer(Trainer):
    """utils/trainer.py:571-596 -- pretrained (frozen) adaptive front end + separator trained on its
    output (BASELINE config 2).  model_folder restore is replaced by an optional state dict."""

    def __init__(self, separator, name="Front_Separator", front_state=None, **kwargs):
        self.separator_class, self.name, self.front_state = separator, name, front_state
        super().__init__(**kwargs)

    def build(self):
        args = {k: v for k, v in self.args.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        args["pretraining"] = False
        self.model = Adapt(**args)
        self.sepNet = self.model.connect_front(self.separator_class)
        self.store = self.model.store

    def post_build(self):
        if self.front_state:
            self.store.load_state_dict(self.front_state, strict=False)
        self.model.freeze_all_with("front/")
        self.model.freeze_all_with("back/")

    def loss(self, x_mix, x_non_mix, ind):
        B = x_mix.shape[0]
        with torch.no_grad():
            y, _ = self.model.front(x_mix, x_non_mix)
        inp = self.sepNet.plugged_inputs(y, B)
        V = self.sepNet.prediction(inp["X"].contiguous())
        return self.sepNet.cost(V, inp["labels"], ind)


class Adapt_Pretrainer(Trainer):
    """utils/trainer.py:529-537 -- raw-waveform autoencoder pre-training of the adaptive front/back end
    (BASELINE config 4: `--loss sdr+l2 --separation mask --beta 0.01`, README.md:23)."""

    def __init__(self, name="AdaptiveNet", **kwargs):
        self.name = name
        super().__init__(**kwargs)

    def build(self):
        args = {k: v for k, v in self.args.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        args["pretraining"] = True
        self.model = Adapt(**args)
        self.store = self.model.store

    def loss(self, x_mix, x_non_mix, ind):
        cost, self.aux = self.model.cost(x_mix, x_non_mix)
        return cost


class STFT_Separator_enhance_Trainer(Trainer):
    """utils/trainer.py:488-500 -- a trained STFT separator (frozen) + the enhance BLSTM layer trained on the
    k-means separated magnitudes with the PIT-L2 enhance cost (BASELINE config 3, second stage).
    init_idx (optional): the k-means initial rows, otherwise drawn like the reference (np.random.choice)."""

    def __init__(self, separator, name="STFT_Separator_enhance", separator_state=None, **kwargs):
        self.separator_class, self.name, self.separator_state = separator, name, separator_state
        self.init_idx = None
        super().__init__(**kwargs)

    def build(self):
        args = {k: v for k, v in self.args.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        self.model = self.separator_class(plugged=False, **args)
        self.model.add_enhance_layer()
        self.store = self.model.store

    def post_build(self):
        if self.separator_state:
            self.store.load_state_dict(self.separator_state, strict=False)
        self.model.freeze_all_except("enhance/")

    def loss(self, x_mix, x_non_mix, ind):
        m = self.model
        pre = m.preprocessing(x_mix, x_non_mix, want_mag_non_mix=True)
        with torch.no_grad():                       # hard k-means labels are not differentiable (network.py:554-582)
            V = m.prediction(pre["X"])
            sep, _ = m.separate(V, pre["X"], self.init_idx)
        _, cost_in = m.enhance(sep, pre["X"])
        return m.enhance_cost(cost_in, pre["X_non_mix"])


class STFT_Separator_FineTune_Trainer(Trainer):
    """utils/trainer.py:502-526 -- end-to-end fine-tuning of the STFT pipeline: |STFT| -> separator (k-means masks) ->
    enhance BLSTM layer -> postprocessing (mixture phase, inverse STFT) -> PIT waveform loss (`cost_finetuning`,
    models/network.py:697-723).  Trains the variables whose name contains one of `train` (--train substrings); the
    gradient reaches them through the enhance layer's mask output (differentiable inverse STFT); hard k-means labels and
    the separator trunk they come from carry no gradient.  The inverse STFT returns (T-1)*hop + frame samples: the
    targets are cropped to that length, as TF's inverse_stft does."""

    def __init__(self, separator, name="STFT_Separator_FineTune", state=None, train=("enhance",), **kwargs):
        self.separator_class, self.name, self.state, self.train_names = separator, name, state, tuple(train)
        self.init_idx = None
        super().__init__(**kwargs)

    def build(self):
        args = {k: v for k, v in self.args.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        self.model = self.separator_class(plugged=False, **args)
        self.model.add_enhance_layer()
        self.store = self.model.store

    def post_build(self):
        if self.state:
            self.store.load_state_dict(self.state, strict=False)
        self.store.set_trainable(lambda name: any(t in name for t in self.train_names))

    def loss(self, x_mix, x_non_mix, ind):
        m = self.model
        spec, X = ops.stft(x_mix.contiguous(), m.window_size, m.hop_size)
        with torch.no_grad():
            V = m.prediction(X)
            sep, _ = m.separate(V, X, self.init_idx)
        m.enhance(sep, X)
        out = m.postprocessing_masks(spec, m.enhance_masks)                   # [B,S,L']
        return m.cost_finetuning(x_non_mix[:, :, :out.shape[2]], out)


class Front_Separator_Enhance_Finetuning_Trainer(Trainer):
    """utils/trainer.py:636-658 -- end-to-end fine-tuning of the adaptive pipeline: front -> separator (k-means masks)
    -> enhance BLSTM layer -> back (unpool + transposed conv) -> PIT waveform loss (`cost_finetuning`,
    models/adapt.py:404-431).  Only the variables whose name contains one of `train` (the reference's --train
    substrings, utils/trainer.py:119-120, :648-653) are optimised; model_folder restore is replaced by an optional
    state dict.  The hard k-means labels are not differentiable (network.py:554-582): the gradient reaches the
    enhance layer, the back filterbank and -- through the masked front output X * mask -- the front filterbank.
    init_idx (optional): the k-means initial rows, otherwise drawn like the reference (np.random.choice)."""

    def __init__(self, separator, name="Front_Separator_Enhance_Finetuning", state=None, train=("enhance", "back"), **kwargs):
        self.separator_class, self.name, self.state, self.train_names = separator, name, state, tuple(train)
        self.init_idx = None
        super().__init__(**kwargs)

    def build(self):
        args = {k: v for k, v in self.args.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        args["pretraining"] = False
        self.model = Adapt(**args)
        self.sepNet = self.model.connect_front(self.separator_class)
        self.sepNet.add_enhance_layer()
        self.store = self.model.store

    def post_build(self):
        if self.state:
            self.store.load_state_dict(self.state, strict=False)
        self.store.set_trainable(lambda name: any(t in name for t in self.train_names))

    def loss(self, x_mix, x_non_mix, ind):
        m, sn = self.model, self.sepNet
        B, Lw = x_mix.shape
        y, am = m.front(x_mix, x_non_mix)
        X = y[:B].contiguous()
        with torch.no_grad():                       # hard k-means labels are not differentiable
            V = sn.prediction(X.detach())
            _, lab = sn.separate(V, X.detach(), self.init_idx)
        Xf = X.reshape(B, -1)
        sep = torch.stack([Xf * (lab == k).to(Xf.dtype) for k in range(sn.S)], 1).reshape(B * sn.S, X.shape[1], X.shape[2]) \
            if sn.beta is None else (Xf.unsqueeze(1) * lab.transpose(1, 2)).reshape(B * sn.S, X.shape[1], X.shape[2])
        enhanced, _ = sn.enhance(sep, X)                                       # [B,S,TF]
        back = m.back(enhanced.reshape(B * sn.S, X.shape[1], X.shape[2]), am, B, Lw)
        return m.cost_finetuning(x_non_mix, back)


class STFT_Separator_Inference:
    """utils/trainer.py:406-417 + Trainer.inference (:190-229): mixture -> separated waveforms."""

    def __init__(self, separator, **kwargs):
        args = {k: v for k, v in kwargs.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        self.model = separator(plugged=False, **args)
        self.model.finalize()

    @torch.no_grad()
    def infer(self, x_mix, init_idx=None):
        m = self.model
        spec, X = ops.stft(x_mix.contiguous(), m.window_size, m.hop_size)
        V = m.prediction(X)
        _, lab = m.separate(V, X, init_idx)
        return m.postprocessing(spec, lab)


class Front_Separator_Inference:
    """utils/trainer.py:420-434: front -> separator k-means masks -> back (unpool + transposed conv)."""

    def __init__(self, separator, **kwargs):
        args = {k: v for k, v in kwargs.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        args["pretraining"] = False
        self.model = Adapt(**args)
        self.sepNet = self.model.connect_front(separator)
        self.model.finalize()

    @torch.no_grad()
    def infer(self, x_mix, x_non_mix, init_idx=None):
        B, Lw = x_mix.shape
        y, am = self.model.front(x_mix, x_non_mix)
        X = y[:B].contiguous()
        V = self.sepNet.prediction(X)
        sep, _ = self.sepNet.separate(V, X, init_idx)
        return self.model.back(sep, am, B, Lw)
