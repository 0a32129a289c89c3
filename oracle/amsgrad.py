"""Restatement of the AMSGrad optimizer (utils/ops.py:639-704, 781-792) as used by
Network.optimize (models/network.py:181-182: AMSGrad(lr, beta1=0.9, beta2=0.99,
epsilon=1e-3) with a constant learning rate) plus tf.clip_by_global_norm
(models/network.py:191-192).  Test infrastructure (see oracle/__init__.py)."""
import math

import torch


class AMSGrad:
    def __init__(self, params, lr, beta1=0.9, beta2=0.99, eps=1e-3, clip=0.0):
        self.params = params                      # dict name -> tensor (updated in place)
        self.lr, self.b1, self.b2, self.eps, self.clip = lr, beta1, beta2, eps, clip
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.vhat = {k: torch.zeros_like(v) for k, v in params.items()}
        self.b1p, self.b2p = beta1, beta2         # beta powers start at beta (utils/ops.py:668-669)

    def increment_epoch(self):
        pass                                      # the 'Adam' branch ignores the decayed rate (models/network.py:181-182)

    def step(self, grads):
        grads = _clipped(grads, self.clip)
        lr_t = self.lr * math.sqrt(1 - self.b2p) / (1 - self.b1p)
        with torch.no_grad():
            for k, g in grads.items():
                self.m[k].mul_(self.b1).add_(g * (1 - self.b1))
                self.v[k].mul_(self.b2).add_(g * g * (1 - self.b2))
                torch.maximum(self.v[k], self.vhat[k], out=self.vhat[k])
                self.params[k].sub_(lr_t * self.m[k] / (torch.sqrt(self.vhat[k]) + self.eps))
        self.b1p *= self.b1
        self.b2p *= self.b2


def exponential_decay(lr, global_epoch, decay_epoch, rate=0.5):
    """tf.train.exponential_decay(lr, global_epoch, decay_epoch, 0.5, staircase=True) (models/network.py:175-177):
    lr * rate ** floor(global_epoch / decay_epoch)."""
    return lr * rate ** (global_epoch // decay_epoch)


def _clipped(grads, clip):
    """tf.clip_by_global_norm (models/network.py:191-192)."""
    if clip == 0.0:
        return grads
    gn = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    scale = clip / max(gn, clip)
    return {k: g * scale for k, g in grads.items()}


class Momentum:
    """tf.train.MomentumOptimizer(learning_rate, momentum=0.9) as --optimizer SGD selects it (models/network.py:183):
    accum = momentum * accum + g ; var -= lr * accum (use_nesterov=False), lr decayed per epoch (:175-177)."""

    def __init__(self, params, lr, momentum=0.9, decay_epoch=50, clip=0.0):
        self.params, self.lr, self.momentum, self.decay_epoch, self.clip = params, lr, momentum, decay_epoch, clip
        self.accum = {k: torch.zeros_like(v) for k, v in params.items()}
        self.global_epoch = 0

    def increment_epoch(self):
        self.global_epoch += 1

    def step(self, grads):
        grads = _clipped(grads, self.clip)
        lr = exponential_decay(self.lr, self.global_epoch, self.decay_epoch)
        with torch.no_grad():
            for k, g in grads.items():
                self.accum[k].mul_(self.momentum).add_(g)
                self.params[k].sub_(lr * self.accum[k])


class RMSProp:
    """tf.train.RMSPropOptimizer(learning_rate) as --optimizer RMSProp selects it (models/network.py:185), TF 1.x defaults:
    decay 0.9, momentum 0.0, epsilon 1e-10, centered False; slots: rms initialised to ONES, momentum to zeros;
    ms = decay * ms + (1 - decay) * g^2 ; mom = momentum * mom + lr * g / sqrt(ms + epsilon) ; var -= mom."""

    def __init__(self, params, lr, decay=0.9, momentum=0.0, epsilon=1e-10, decay_epoch=50, clip=0.0):
        self.params, self.lr, self.decay, self.momentum, self.eps = params, lr, decay, momentum, epsilon
        self.decay_epoch, self.clip = decay_epoch, clip
        self.ms = {k: torch.ones_like(v) for k, v in params.items()}
        self.mom = {k: torch.zeros_like(v) for k, v in params.items()}
        self.global_epoch = 0

    def increment_epoch(self):
        self.global_epoch += 1

    def step(self, grads):
        grads = _clipped(grads, self.clip)
        lr = exponential_decay(self.lr, self.global_epoch, self.decay_epoch)
        with torch.no_grad():
            for k, g in grads.items():
                self.ms[k].mul_(self.decay).add_(g * g * (1 - self.decay))
                self.mom[k].mul_(self.momentum).add_(lr * g / torch.sqrt(self.ms[k] + self.eps))
                self.params[k].sub_(self.mom[k])


def make_optimizer(kind, params, lr, decay_epoch=50, clip=0.0):
    """Network.optimize's switch on --optimizer (models/network.py:181-186)."""
    if kind == "Adam":
        return AMSGrad(params, lr, clip=clip)
    if kind == "SGD":
        return Momentum(params, lr, decay_epoch=decay_epoch, clip=clip)
    if kind == "RMSProp":
        return RMSProp(params, lr, decay_epoch=decay_epoch, clip=clip)
    raise ValueError(kind)
