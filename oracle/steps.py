"""Whole training steps of the reference graphs, assembled from the restated pieces (test
infrastructure, see oracle/__init__.py): what ONE `sess.run([optimize, cost])` computes
(models/network.py:228-232) for
  * STFT + DPCL / L41                (utils/trainer.py:468-486, BASELINE config 1 / 3 trunk)
  * frozen adaptive front + DPCL     (utils/trainer.py:571-596, BASELINE config 2)
Used by the end-to-end parity tests and as the timed CPU baseline of bench.py
(`cpu_baseline.kind = "port"`: torch-CPU restatement of the TF graph; TF 1.x is not installable)."""
import torch

from . import models as M
from .amsgrad import make_optimizer


def trainable(params, prefixes):
    return {k: v for k, v in params.items() if any(k.startswith(p) for p in prefixes)}


def stft_separator_loss(p, x_mix, x_non_mix, I, *, nb_layers, embedding_size, window_size=512, hop_size=256,
                        loss="dpcl", normalize=True):
    a, b = (1.0, 0.0) if loss == "dpcl" else (1.0, -1.0)
    pre = M.separator_preprocessing(x_mix, x_non_mix, window_size, hop_size, a, b)
    V = M.separator_prediction(p, pre["X"], nb_layers, embedding_size, normalize)
    cost = M.dpcl_cost(V, pre["y"]) if loss == "dpcl" else M.l41_cost(p, V, pre["y"], I, normalize)
    return cost, {"V": V, "pre": pre}


def front_separator_loss(p, x_mix, x_non_mix, I, *, nb_layers, embedding_size, max_pool, hop, loss="dpcl",
                         normalize=True):
    a, b = (1.0, 0.0) if loss == "dpcl" else (1.0, -1.0)
    B, S, L = x_non_mix.shape
    with torch.no_grad():
        fr = M.adapt_front(p, x_mix, x_non_mix, max_pool, hop, True)
    inp = M.separator_plugged_inputs(fr["y"], B, S, a, b)
    V = M.separator_prediction(p, inp["X"], nb_layers, embedding_size, normalize)
    cost = M.dpcl_cost(V, inp["y"]) if loss == "dpcl" else M.l41_cost(p, V, inp["y"], I, normalize)
    return cost, {"V": V, "front": fr, "inp": inp}


class Stepper:
    """fwd + bwd + optimizer (AMSGrad by default) on the parameters whose names start with `train_prefixes`."""

    def __init__(self, params, loss_fn, train_prefixes=("prediction/", "speaker_centroids"), lr=1e-3, clip=0.0,
                 optimizer="Adam", decay_epoch=50):
        self.p = params
        self.loss_fn = loss_fn
        self.tr = trainable(params, train_prefixes)
        for v in self.tr.values():
            v.requires_grad_(True)
        self.opt = make_optimizer(optimizer, self.tr, lr, decay_epoch=decay_epoch, clip=clip)

    def step(self, x_mix, x_non_mix, I):
        cost, aux = self.loss_fn(self.p, x_mix, x_non_mix, I)
        grads = torch.autograd.grad(cost, list(self.tr.values()), allow_unused=True)
        grads = {k: (g if g is not None else torch.zeros_like(v)) for (k, v), g in zip(self.tr.items(), grads)}
        self.opt.step(grads)
        self.last_grads = grads
        return float(cost.detach()), aux


def stft_enhance_loss(p, x_mix, x_non_mix, I, *, nb_layers, embedding_size, window_size=512, hop_size=256,
                      nb_layers_enhance=3, nb_tries=10, nb_steps=10, init_idx=None, seed=0):
    """STFT_Separator_enhance_Trainer's step (utils/trainer.py:488-500, BASELINE config 3): a trained separator (no gradient)
    -> k-means masks (models/network.py:554-582) -> enhance BLSTM layer (:610-660) -> PIT-L2 enhance cost (:662-693)."""
    import numpy as np
    from .kmeans import KMeans, random_init_idx
    S = x_non_mix.shape[1]
    pre = M.separator_preprocessing(x_mix, x_non_mix, window_size, hop_size, 1.0, -1.0)
    with torch.no_grad():
        V = M.separator_prediction(p, pre["X"], nb_layers, embedding_size, True)
        B, Tt, Fb, E = V.shape
        if init_idx is None:
            init_idx = random_init_idx(B * nb_tries, Tt * Fb, S, np.random.RandomState(seed))
        km = KMeans(nb_clusters=S, nb_tries=nb_tries, nb_iterations=nb_steps)
        sep, masks = M.separate(V, pre["X"], lambda emb: km.fit(emb, init_idx=init_idx)[1], S)
    _, cost_in, _ = M.enhance(p, sep, pre["X"], S, nb_layers_enhance)
    cost = M.enhance_cost(cost_in, pre["X_non_mix"])
    return cost, {"V": V, "pre": pre, "masks": masks}


def stft_inference(p, x_mix, x_non_mix, I, *, nb_layers, embedding_size, window_size=512, hop_size=256, S=2, nb_tries=10,
                   nb_steps=10, init_idx=None, seed=0):
    """STFT_Separator_Inference (utils/trainer.py:406-417; BASELINE config 5): |STFT| -> embeddings -> k-means masks ->
    postprocessing (mixture phase, inverse STFT) -> separated waveforms [B,S,L']."""
    import numpy as np
    from .kmeans import KMeans, random_init_idx
    stfts = T_stft(x_mix, window_size, hop_size)
    X = stfts.abs()
    V = M.separator_prediction(p, X, nb_layers, embedding_size, True)
    B, Tt, Fb, E = V.shape
    if init_idx is None:
        init_idx = random_init_idx(B * nb_tries, Tt * Fb, S, np.random.RandomState(seed))
    km = KMeans(nb_clusters=S, nb_tries=nb_tries, nb_iterations=nb_steps)
    sep, masks = M.separate(V, X, lambda emb: km.fit(emb, init_idx=init_idx)[1], S)
    out = M.postprocessing(sep, stfts, S, window_size, hop_size)
    return out, {"V": V, "masks": masks}


def T_stft(x, window_size, hop_size):
    from . import tf_ops
    return tf_ops.stft(x, window_size, hop_size)
