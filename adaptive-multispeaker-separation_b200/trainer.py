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


class Optimizer:
    """Network.optimize (models/network.py:167-194) over the flat parameter buffer, one fused kernel per contiguous
    trainable segment (normally one):
      * 'Adam'    -> AMSGrad(lr, beta1=0.9, beta2=0.99, epsilon=1e-3), CONSTANT learning rate (:181-182, utils/ops.py:639-704);
      * 'SGD'     -> tf.train.MomentumOptimizer(decayed lr, momentum=0.9) (:183);
      * 'RMSProp' -> tf.train.RMSPropOptimizer(decayed lr) with TF 1.x's defaults decay 0.9, momentum 0, epsilon 1e-10 and
                     the rms slot initialised to ONE (:185);
      * decayed lr = exponential_decay(lr, global_epoch, decay_epoch, 0.5, staircase=True) = lr * 0.5^(epoch // decay_epoch)
                     (:175-177); global_epoch advances through increment_epoch() once per epoch (utils/trainer.py:353);
      * optional tf.clip_by_global_norm on the gradient of the batch-MEAN loss (:191-192): the norm is taken of
        grad_scale * g, because under data parallelism the buffer holds the sum over ranks and grad_scale = 1 / world size.
    Variables frozen with set_trainable() (the reference removes them from var_list) are not in any segment: they are
    neither updated nor do they keep stale momentum."""

    KINDS = ("Adam", "SGD", "RMSProp")

    def __init__(self, store, lr, kind="Adam", decay_epoch=50, clip=0.0, beta1=0.9, beta2=0.99, eps=1e-3):
        if kind not in self.KINDS:
            raise ValueError(f"--optimizer {kind}: the reference knows {self.KINDS} (models/network.py:181-186)")
        self.store, self.kind, self.lr, self.decay_epoch, self.clip = store, kind, float(lr), int(decay_epoch), float(clip or 0.0)
        self.b1, self.b2, self.eps = beta1, beta2, eps
        n = store.n_trainable
        z = lambda: torch.zeros(n, dtype=torch.float32, device=store.device)  # noqa: E731
        if kind == "Adam":
            self.m, self.v, self.vhat = z(), z(), z()
        elif kind == "SGD":
            self.accum = z()
        else:
            self.ms, self.mom = torch.ones(n, dtype=torch.float32, device=store.device), z()
        self.t = 0
        self.global_epoch = 0

    def increment_epoch(self):
        self.global_epoch += 1

    def learning_rate(self):
        """The rate the next step uses (tf.train.exponential_decay, staircase; 'Adam' ignores the decay)."""
        if self.kind == "Adam":
            return self.lr
        return self.lr * 0.5 ** (self.global_epoch // self.decay_epoch)

    def state_tensors(self):
        return {"Adam": ("m", "v", "vhat"), "SGD": ("accum",), "RMSProp": ("ms", "mom")}[self.kind]

    def step(self, grad_scale=1.0):
        st = self.store
        self.t += 1
        segs = st.trainable_segments()
        fac = None
        if self.clip:
            fac = ops.global_norm_clip_factor([st.grad_flat[o:o + n] for o, n in segs], self.clip, grad_scale)
        for o, n in segs:
            p, g = st.flat[o:o + n], st.grad_flat[o:o + n]
            if self.kind == "Adam":
                lr_t = ops.amsgrad_lr_t(self.lr, self.b1, self.b2, self.t)
                ops.amsgrad_step(p, g, self.m[o:o + n], self.v[o:o + n], self.vhat[o:o + n], lr_t, self.b1, self.b2,
                                 self.eps, grad_scale, fac)
            elif self.kind == "SGD":
                ops.momentum_step(p, g, self.accum[o:o + n], self.learning_rate(), 0.9, grad_scale, fac)
            else:
                ops.rmsprop_step(p, g, self.ms[o:o + n], self.mom[o:o + n], self.learning_rate(), 0.9, 0.0, 1e-10,
                                 grad_scale, fac)


class AMSGradOptimizer(Optimizer):
    """Kept for callers of the first round: Optimizer(kind='Adam')."""

    def __init__(self, store, lr, beta1=0.9, beta2=0.99, eps=1e-3, clip=0.0):
        super().__init__(store, lr, "Adam", clip=clip, beta1=beta1, beta2=beta2, eps=eps)


class DevicePrefetcher:
    """Host -> device input staging one step ahead (the role of the reference's tf.data prefetch(1),
    data/dataset.py:506-514): pinned host batches are copied on a side stream into double-buffered device
    inputs while the previous step computes; stream events (no host sync) order copy and compute."""

    def __init__(self, example_batch):
        self.stream = torch.cuda.Stream()
        self.bufs = [[None if a is None else torch.empty(tuple(a.shape), dtype=a.dtype, device="cuda") for a in example_batch]
                     for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [None, None]
        self.k = 0

    @staticmethod
    def pin(batch):
        return tuple(None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in batch)

    def submit(self, pinned_batch):
        """Start the H2D copy of a pinned batch; returns the slot to pass to get()."""
        k = self.k
        self.k ^= 1
        if self.done[k] is not None:
            self.stream.wait_event(self.done[k])          # the step that last read this slot has finished
        with torch.cuda.stream(self.stream):
            for d, h in zip(self.bufs[k], pinned_batch):
                if d is not None:
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
        self.optimizer = Optimizer(self.store, self.args.get("learning_rate", DEFAULTS["learning_rate"]),
                                   kind=self.args.get("optimizer", DEFAULTS["optimizer"]),
                                   decay_epoch=self.args.get("decay_epoch", DEFAULTS["decay_epoch"]),
                                   clip=self.args.get("gradient_norm_clip", 0.0))

    # -- to override ---------------------------------------------------------------------------
    def build(self):
        raise NotImplementedError

    def post_build(self):
        pass

    def loss(self, x_mix, x_non_mix, ind):
        raise NotImplementedError

    # -- optional CUDA-graph replay of forward + backward ---------------------------------------
    def enable_cuda_graph(self, eager_steps=2):
        """Capture zero-grad + loss + backward of one step into a CUDA graph (after `eager_steps` ordinary steps, which
        also warm every lazily initialised kernel attribute and the NCCL communicator) and replay it from then on; under
        data parallelism the per-layer gradient all-reduces are part of the graph (launched from the backward pass as each
        layer's gradients complete); the optimizer stays outside (its step count changes per step).  Each train_step() call is still exactly
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
        cg["inputs"] = [None if t is None else t.clone() for t in (x_mix, x_non_mix, ind)]
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            self.store.grad_flat.zero_()
            cost = self.loss(*self.prepare(cg["inputs"][0], cg["inputs"][1]), cg["inputs"][2])
            cg["scale"] = self._backward(cost)           # under data parallelism the NCCL kernels are captured too
        cg["graph"], cg["cost"] = graph, cost.detach()
        cg["kernels"] = _lib.launch_count() - n0          # library kernels recorded in the graph = run by every replay
        cg["replays"] = 0

    # -- the input contract on the device ---------------------------------------------------------
    def prepare(self, x_mix, x_non_mix):
        """(x_mix or None, x_non_mix) -> (x_mix, x_non_mix) as the graph expects them (models/network.py:44-88): with
        --dataset_normalize every source is normalised to zero mean / unit variance and the mixture is rebuilt from the
        normalised sources (data/dataset.py:456-468; the reference normalises whole utterances before chunking, here the
        unit is the chunk the caller hands over -- the data layer itself is out of scope); x_mix=None builds the
        mixture on the device as the sum of the sources, so the host ships a third less."""
        if self.args.get("dataset_normalize", False):
            x_non_mix = x_non_mix.clone()
            x_mix, self.norm_stats = ops.prepare_inputs(x_non_mix, normalize=True)
        elif x_mix is None:
            x_mix, _ = ops.prepare_inputs(x_non_mix.contiguous())
        return x_mix, x_non_mix

    # -- the gradient exchange: per-layer buckets launched from the backward pass (dp.GradBuckets) ---------------------
    def _buckets(self):
        """(Re)build the buckets and the post-accumulate-grad hooks when the trainable set changed."""
        st = self.store
        if getattr(self, "_gb_key", None) is not st.trainable_segments():
            for h in getattr(self, "_gb_hooks", []):
                h.remove()
            groups = st.grad_buckets()
            self._gb = dp.GradBuckets(st.grad_flat, [(lo, hi, len(names)) for lo, hi, names in groups])
            self._gb_hooks = []
            for i, (_, _, names) in enumerate(groups):
                for nme in names:
                    self._gb_hooks.append(st[nme].register_post_accumulate_grad_hook(lambda p_, i=i: self._gb.ready(i)))
            self._gb_key = st.trainable_segments()
        return self._gb

    def _backward(self, cost):
        """backward + the gradient exchange; returns the scale (1 / world size) the optimizer applies.  Single process:
        plain backward, no hooks consulted."""
        if not self.distributed or not self.args.get("overlap_allreduce", True):
            cost.backward()
            lo, hi = self.store.trainable_span()
            return dp.allreduce_sum_(self.store.grad_flat[lo:hi])    # the single collective of the path
        gb = self._buckets()
        gb.reset()
        cost.backward()
        return gb.finish()

    # -- one optimisation step on device tensors ------------------------------------------------
    def train_step(self, x_mix, x_non_mix, ind):
        """Network.train (models/network.py:228-232): forward + backward + optimizer on one batch; returns the cost
        (device scalar).  x_mix may be None (built on the device from the sources)."""
        cg = getattr(self, "_cg", None)
        if cg is not None:
            if cg["graph"] is None and cg["calls"] >= cg["eager"]:
                self._capture_step(x_mix, x_non_mix, ind)
            cg["calls"] += 1
            if cg["graph"] is not None:
                for dst, src in zip(cg["inputs"], (x_mix, x_non_mix, ind)):
                    if dst is not None:
                        dst.copy_(src, non_blocking=True)
                cg["graph"].replay()                     # forward + backward + the bucketed gradient all-reduce
                cg["replays"] += 1
                self.optimizer.step(cg["scale"])
                return cg["cost"].clone()
        self.store.grad_flat.zero_()
        cost = self.loss(*self.prepare(x_mix, x_non_mix), ind)
        scale = self._backward(cost)
        self.optimizer.step(scale)
        return cost.detach()

    @torch.no_grad()
    def eval_step(self, x_mix, x_non_mix, ind):
        """Network.valid_batch / test_batch (models/network.py:240-262): the cost of one batch, no update."""
        return self.loss(*self.prepare(x_mix, x_non_mix), ind).detach()

    def run_steps(self, data, steps, log_every=0):
        """data: iterator of host batches (numpy or pinned tensors; the mixture entry may be None).  Inputs are staged one
        step ahead (DevicePrefetcher) and each step's cost is read back while the next step runs.
        Returns the list of per-step costs (floats)."""
        costs = []
        t0 = time.time()
        as_pinned = lambda b: b if _is_pinned(b) else DevicePrefetcher.pin(b)  # noqa: E731
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

    # -- the reference's training loop ------------------------------------------------------------
    def _mean_cost(self, batches):
        """Mean cost over an iterable of host batches (utils/trainer.py:329-335); under data parallelism every rank
        evaluates its own batches and the per-rank means are averaged (a scalar exchange outside the training step)."""
        costs = []
        for hb in batches:
            dev = [None if a is None else torch.as_tensor(np.ascontiguousarray(a)).cuda() for a in hb]
            costs.append(self.eval_step(*dev))
        if not costs:
            return float("nan")
        c = torch.stack(costs).mean().reshape(1)
        if self.distributed:
            torch.distributed.all_reduce(c)
            c /= self.world
        return float(c)

    def save(self, step):
        """Network.save (models/network.py:223-226): <log_dir>/<name>/<runID>/model-<step>/{params, model.npz}."""
        path = os.path.join(self.log_dir, getattr(self, "name", "model"), self.runID, f"model-{step}")
        if self.rank == 0:
            net = getattr(self, "model", None)
            net.save(path)
        if self.distributed:
            torch.distributed.barrier()
        return path

    def train(self, data, steps=None, log_every=0, log_dir=None, runID=None, verbose=True):
        """Trainer.train (utils/trainer.py:264-390).  `data` is either
          * a dataset object with `train()`, `valid()`, `test()` methods, each returning a fresh iterable of host batches
            (mix or None, non_mix, ind) -- the role of TFDataset's three initialisable iterators (data/dataset.py:520-645).
            Then the reference's loop runs: `epochs` passes over train(); every `validation_step` steps the mean cost over
            valid() and a checkpoint if it improved (:323-346); increment_epoch per epoch (:353: drives the SGD / RMSProp
            learning-rate decay); after the last epoch one more validation + save-if-best (:358-375), the best
            checkpoint is restored and the mean cost over test() reported (:380-388).  Returns a dict with the history;
          * or an iterator of host batches together with `steps`: a plain step loop (run_steps), returns the costs."""
        if steps is not None or not hasattr(data, "train"):
            return self.run_steps(iter(data), steps, log_every)
        import tempfile
        self.log_dir = log_dir or self.args.get("log_dir") or tempfile.mkdtemp(prefix="amss_log_")
        self.runID = runID or self.args.get("runID") or time.strftime("run-%Y%m%d-%H%M%S")
        nb_epochs = int(self.args.get("epochs", DEFAULTS["epochs"]))
        vstep = int(self.args.get("validation_step", DEFAULTS["validation_step"]))
        say = print if (verbose and self.rank == 0) else (lambda *a, **k: None)
        best, best_path, step = 1e100, "", 0
        hist = {"train_costs": [], "valid": [], "saved": [], "learning_rates": []}
        time_spent, t1 = [0.0] * 10, time.time()

        def validate():
            nonlocal best, best_path
            t = time.time()
            vc = self._mean_cost(data.valid())
            hist["valid"].append((step, vc))
            if vc < best:                                    # save the model if it is better (:338-342)
                best = vc
                best_path = self.save(step)
                hist["saved"].append((step, best_path))
                say("Save best model with :", best)
            say(f"Validation set tested in {time.time() - t:.3f} seconds\nValidation set:  {vc}")

        pf = None
        for epoch in range(nb_epochs):
            hist["learning_rates"].append(self.optimizer.learning_rate())
            src = data.train()                                        # training_initializer (:313)
            nb = len(src) if hasattr(src, "__len__") else getattr(data, "nb_batches_train", None)
            it = iter(src)
            nxt = next(it, None)
            slot = None
            if nxt is not None:
                nxt = nxt if _is_pinned(nxt) else DevicePrefetcher.pin(nxt)
                pf = pf or DevicePrefetcher(nxt)
                slot = pf.submit(nxt)
            b = 0
            while slot is not None:
                cost_dev = self.train_step(*pf.get(slot))
                pf.release(slot)
                nxt = next(it, None)                                  # stage the next batch while this step runs
                slot = None
                if nxt is not None:
                    nxt = nxt if _is_pinned(nxt) else DevicePrefetcher.pin(nxt)
                    slot = pf.submit(nxt)
                c = float(cost_dev)
                if c != c:
                    raise FloatingPointError(f"NaN cost at step {step} (epoch {epoch + 1}, batch {b + 1})")
                hist["train_costs"].append(c)
                if (step + 1) % vstep == 0:
                    validate()
                time_spent = time_spent[1:] + [time.time() - t1]      # running mean over the last 10 steps (:348-351)
                avg = sum(time_spent) / len(time_spent)
                eta = f"{avg * ((nb_epochs - epoch - 1) * nb + (nb - b - 1)):.0f} s" if nb else "?"
                say(f"Epoch # {epoch + 1} / {nb_epochs}  Batch # {b + 1} / {nb or '?'} in {avg:.4f} sec loss= {c}  ETA = {eta}")
                t1 = time.time()
                step += 1
                b += 1
            self.optimizer.increment_epoch()                          # sess.run(increment_epoch) (:353)
        validate()                                                    # validation at the last step (:358-375)
        say("Best model with Validation:  ", best, "\nPath = ", best_path)
        if best_path:                                                 # restore_last_checkpoint (:380)
            self.model.restore_model(best_path)
        test_cost = self._mean_cost(data.test())
        say("Test cost = ", test_cost)
        hist.update(best_validation_cost=best, best_path=best_path, test_cost=test_cost, steps=step)
        return hist


def _is_pinned(batch):
    t = next((a for a in batch if a is not None), None)
    return torch.is_tensor(t) and t.is_pinned()


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
        if inp["weights"] is not None:
            return self.sepNet.cost(V, inp["labels"], ind, inp["weights"])
        return self.sepNet.cost(V, inp["labels"], ind)


class Front_Separator_Enhance_Trainer(Trainer):
    """utils/trainer.py:600-610 + Adapt.connect_enhance_to_separator (models/adapt.py:456-469): pretrained front / back and a
    trained separator, all frozen; the enhance BLSTM layer is trained on the k-means separated front output with the PIT-L2
    enhance cost against the sources' front responses (models/network.py:662-693 with the plugged X_non_mix).
    model_folder restore is replaced by an optional state dict; init_idx as in STFT_Separator_enhance_Trainer."""

    def __init__(self, separator, name="Front_Separator_Enhance", state=None, **kwargs):
        self.separator_class, self.name, self.state = separator, name, state
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
        self.model.freeze_all_except("enhance/")

    def loss(self, x_mix, x_non_mix, ind):
        sn = self.sepNet
        B = x_mix.shape[0]
        with torch.no_grad():
            y, _ = self.model.front(x_mix, x_non_mix)
            inp = sn.plugged_inputs(y, B)
            X_in = inp["X_raw"].contiguous()
            V = sn.prediction(inp["X"].contiguous())
            sep, _ = sn.separate(V, X_in, self.init_idx)
            Tp, N = y.shape[1], y.shape[2]
            X_non_mix = y[B:].reshape(B, sn.S, Tp, N).permute(0, 2, 3, 1).contiguous()          # [B,T,N,S] (network.py:372)
        return sn.enhance_cost_fused(sn.enhance_logits(sep, X_in), X_in, X_non_mix)


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
        cost, aux = self.model.cost(x_mix, x_non_mix)
        # detached: holding the autograd graph of a finished step would keep its AccumulateGrad nodes (and their stream) alive,
        # which breaks the capture of the next step into a CUDA graph
        self.aux = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in aux.items()}
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
            sep, _ = m.separate(V, pre["X_input"], self.init_idx)
        return m.enhance_cost_fused(m.enhance_logits(sep, pre["X_input"]), pre["X_input"], pre["X_non_mix"])


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
            V = m.prediction(m._prep(X, False))
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
            V = sn.prediction(sn._prep(X.detach(), True))
            _, lab = sn.separate(V, X.detach(), self.init_idx)
        Xf = X.reshape(B, -1)
        sep = torch.stack([Xf * (lab == k).to(Xf.dtype) for k in range(sn.S)], 1).reshape(B * sn.S, X.shape[1], X.shape[2]) \
            if sn.beta is None else (Xf.unsqueeze(1) * lab.transpose(1, 2)).reshape(B * sn.S, X.shape[1], X.shape[2])
        enhanced, _ = sn.enhance(sep, X)                                       # [B,S,TF]
        back = m.back(enhanced.reshape(B * sn.S, X.shape[1], X.shape[2]), am, B, Lw)
        return m.cost_finetuning(x_non_mix, back)


class _StreamingInference:
    """Trainer.inference (utils/trainer.py:190-229) as a streaming generator: host batches are staged one step ahead from
    pinned memory (DevicePrefetcher), every batch's separated waveforms [B,S,L'] are copied back into double-buffered
    pinned host memory on a side stream, and the generator yields them one step late (so that the copy overlaps the next
    batch's kernels).  A yielded tensor is valid until two more batches have been yielded.  Replicas only under
    multi-GPU: each rank streams its own batches, no collective."""

    def _infer_batch(self, dev_batch):
        raise NotImplementedError

    def inference(self, data, steps=None):
        """Generator over the separated batches of `data` (pinned host tensors [B,S,L]).  Pipelined: the H2D copy of batch
        n+1 and the D2H copy of batch n-1 overlap the kernels of batch n.  The yielded tensor is one of TWO pinned buffers
        kept on the instance: it is overwritten two batches later (and by the next call) -- copy it to keep it."""
        it = iter(data)
        as_pinned = lambda b: b if _is_pinned(b) else DevicePrefetcher.pin(b)  # noqa: E731
        nxt = next(it, None)
        if nxt is None:
            return
        nxt = as_pinned(nxt)
        pf = DevicePrefetcher(nxt)
        slot = pf.submit(nxt)
        d2h = torch.cuda.Stream()
        # the two pinned result buffers live on the instance: pinning 49 MB costs ~30 ms, once, not once per call
        if not hasattr(self, "_host_out"):
            self._host_out = [None, None]
        host, done, pending, k, n = self._host_out, [None, None], None, 0, 0
        while slot is not None and (steps is None or n < steps):
            # the next batch's H2D copy goes out BEFORE this batch's kernels are queued (it targets the other device slot,
            # last read two steps ago), so it overlaps the whole step
            n += 1
            nxt = next(it, None) if (steps is None or n < steps) else None
            nslot = pf.submit(as_pinned(nxt)) if nxt is not None else None
            out = self._infer_batch(pf.get(slot))
            pf.release(slot)
            slot = nslot
            if host[k] is None or host[k].shape != out.shape:
                host[k] = torch.empty(out.shape, dtype=out.dtype).pin_memory()
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            with torch.cuda.stream(d2h):
                d2h.wait_event(ready)
                host[k].copy_(out, non_blocking=True)
                out.record_stream(d2h)
                done[k] = torch.cuda.Event()
                done[k].record(d2h)
            if pending is not None:
                done[pending].synchronize()
                yield host[pending]
            pending = k
            k ^= 1
        if pending is not None:
            done[pending].synchronize()
            yield host[pending]


class STFT_Separator_Inference(_StreamingInference):
    """utils/trainer.py:406-417 + Trainer.inference (:190-229): mixture -> separated waveforms."""

    def __init__(self, separator, state=None, **kwargs):
        args = {k: v for k, v in kwargs.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        self.model = separator(plugged=False, **args)
        self.model.finalize()
        if state:
            self.model.store.load_state_dict(state, strict=False)
        self.init_idx = None

    @torch.no_grad()
    def infer(self, x_mix, init_idx=None):
        m = self.model
        spec, X = ops.stft(x_mix.contiguous(), m.window_size, m.hop_size)
        V = m.prediction(m._prep(X, False))
        _, lab = m.separate(V, X, init_idx if init_idx is not None else self.init_idx)
        return m.postprocessing(spec, lab)

    def _infer_batch(self, dev_batch):
        return self.infer(dev_batch[0])


class Front_Separator_Inference(_StreamingInference):
    """utils/trainer.py:420-434: front -> separator k-means masks -> back (unpool + transposed conv)."""

    def __init__(self, separator, state=None, **kwargs):
        args = {k: v for k, v in kwargs.items() if k in DEFAULTS or k in ("window_size", "hop_size")}
        args["pretraining"] = False
        self.model = Adapt(**args)
        self.sepNet = self.model.connect_front(separator)
        self.model.finalize()
        if state:
            self.model.store.load_state_dict(state, strict=False)
        self.init_idx = None

    @torch.no_grad()
    def infer(self, x_mix, x_non_mix, init_idx=None):
        B, Lw = x_mix.shape
        y, am = self.model.front(x_mix, x_non_mix)
        X = y[:B].contiguous()
        V = self.sepNet.prediction(self.sepNet._prep(X, True))
        sep, _ = self.sepNet.separate(V, X, init_idx if init_idx is not None else self.init_idx)
        return self.model.back(sep, am, B, Lw)

    def _infer_batch(self, dev_batch):
        x_mix, x_non_mix = dev_batch[0], dev_batch[1]
        if x_mix is None:
            x_mix, _ = ops.prepare_inputs(x_non_mix.contiguous())
        return self.infer(x_mix, x_non_mix)
