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

    def step(self, grads):
        if self.clip != 0.0:
            gn = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
            scale = self.clip / max(gn, self.clip)
            grads = {k: g * scale for k, g in grads.items()}
        lr_t = self.lr * math.sqrt(1 - self.b2p) / (1 - self.b1p)
        with torch.no_grad():
            for k, g in grads.items():
                self.m[k].mul_(self.b1).add_(g * (1 - self.b1))
                self.v[k].mul_(self.b2).add_(g * g * (1 - self.b2))
                torch.maximum(self.v[k], self.vhat[k], out=self.vhat[k])
                self.params[k].sub_(lr_t * self.m[k] / (torch.sqrt(self.vhat[k]) + self.eps))
        self.b1p *= self.b1
        self.b2p *= self.b2
