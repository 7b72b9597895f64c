"""Optimisers (reference optim/optimizer.py:17-196): SGD (momentum / Nesterov), Adagrad, Adadelta, Adam.

The update rules are the reference's array expressions. For fp32 parameters on a cuda device Adam keeps parameters'
gradients and its moments in FLAT buffers (one allocation each): ``zero_grad`` is a flag flip, the whole update is one
multi-tensor kernel launch (pdn_adam_step over the flat segment), and the gradient bucket is the thing the data-parallel
wrapper all-reduces (pydynet_b200/distributed.py) — the 1/world scale is folded into the same kernel.
"""
from math import sqrt

import numpy as np

from ..core import Tensor


class Optimizer:

    def __init__(self, params) -> None:
        self.params = list(params)

    def step(self):
        raise NotImplementedError

    def zero_grad(self):
        for p in self.params:
            p.zero_grad()


class SGD(Optimizer):

    def __init__(self, params, lr: float, momentum: float = .5, weight_decay: float = 0., nesterov=True) -> None:
        super().__init__(params)
        self.lr, self.momentum, self.weight_decay, self.nesterov = lr, momentum, weight_decay, nesterov
        self.v = [_zeros_like(p) for p in self.params]

    def step(self):
        for p, v in zip(self.params, self.v):
            with p.device:
                grad = p.grad + self.weight_decay * p.data
                v *= self.momentum
                v += self.lr * grad
                p.data -= v
                if self.nesterov:
                    p.data -= self.lr * grad


class Adagrad(Optimizer):

    def __init__(self, params, lr: float = 1e-2, weight_decay: float = 0, eps: float = 1e-10) -> None:
        super().__init__(params)
        self.lr, self.weight_decay, self.eps = lr, weight_decay, eps
        self.G = [_zeros_like(p) for p in self.params]

    def step(self):
        for p, G in zip(self.params, self.G):
            with p.device:
                grad = p.grad + self.weight_decay * p.data
                G += grad**2
                p.data -= self.lr * grad / (self.eps + G)**0.5


class Adadelta(Optimizer):

    def __init__(self, params, lr: float = 1.0, rho: float = 0.9, weight_decay: float = 0, eps: float = 1e-6) -> None:
        super().__init__(params)
        self.lr, self.rho, self.eps, self.weight_decay = lr, rho, eps, weight_decay
        self.G = [_zeros_like(p) for p in self.params]

    def step(self):
        for i, p in enumerate(self.params):
            with p.device:
                grad = p.grad + self.weight_decay * p.data
                self.G[i] = self.rho * self.G[i] + (1 - self.rho) * grad**2
                p.data -= self.lr * grad / (self.G[i] + self.eps)**0.5


class Adam(Optimizer):
    """m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps), with one step
    counter ``t`` (starting at 1) shared by all parameters (reference optimizer.py:161-196)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0) -> None:
        super().__init__(params)
        self.lr = lr
        self.beta1, self.beta2 = betas
        self.eps, self.weight_decay = eps, weight_decay
        self.t = 1
        self.grad_scale = 1.0  # set by the data-parallel wrapper to 1/world_size
        self._flat = None
        # flat buffers need one private segment per parameter: parameters that alias one buffer (Parameter(copy=False) tying) keep the
        # reference's per-tensor update, which writes the shared array once per alias
        if self.params and all(p.device.is_cuda and p.dtype == np.float32 and p.requires_grad for p in self.params) \
                and len({p.device for p in self.params}) == 1 and len({p.data.ptr for p in self.params}) == len(self.params):
            from ._flat import FlatAdamState
            self._flat = FlatAdamState(self.params)
            self.m, self.v = self._flat.m_views, self._flat.v_views
        else:
            self.m = [_zeros_like(p) for p in self.params]
            self.v = [_zeros_like(p) for p in self.params]

    def step(self):
        if self._flat is not None:
            from ..cuda import nvtx_range
            with nvtx_range("optimizer.step"):
                self._flat.step(self.lr, self.beta1, self.beta2, self.eps, self.weight_decay, self.t, self.grad_scale)
            self.t += 1
            return
        a_t = sqrt(1 - self.beta2**self.t) / (1 - self.beta1**self.t)
        for p, m, v in zip(self.params, self.m, self.v):
            with p.device:
                grad = p.grad + self.weight_decay * p.data
                m *= self.beta1
                m += (1 - self.beta1) * grad
                v *= self.beta2
                v += (1 - self.beta2) * grad**2
                p.data -= self.lr * a_t * m / (v**0.5 + self.eps)
        self.t += 1


def _zeros_like(p: Tensor):
    with p.device:
        return p.xp.zeros(p.shape, dtype=p.dtype)
