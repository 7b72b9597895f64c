"""Normalisation layers (reference nn/modules/norm.py:9-248).

BatchNorm1d/2d and the reference's ``LayerNorm`` are all the same computation — per-feature statistics over every
*other* axis (the reference LayerNorm reduces the LEADING axes, norm.py:203-208: batch×seq statistics with running
averages used in eval; it is not torch's LayerNorm) — so they share ``_FeatureStatNorm``.  eps defaults to 1e-6,
momentum 0.1, biased variance."""
from .module import Module
from ..parameter import Parameter
from .. import init, functional as F
from ... import core
from ...cuda import Device
from ...special import empty


class _FeatureStatNorm(Module):

    def _setup(self, stat_shape, eps, momentum, device, dtype):
        kw = {"device": Device(device), "dtype": dtype}
        self.eps, self.momentum = eps, momentum
        self.running_mean = Parameter(empty(stat_shape, **kw), requires_grad=False)
        self.running_var = Parameter(empty(stat_shape, **kw), requires_grad=False)
        self.scale = Parameter(empty(stat_shape, **kw))
        self.shift = Parameter(empty(stat_shape, **kw))
        self.reset_parameters()

    def reset_parameters(self):
        init.zeros_(self.running_mean)
        init.ones_(self.running_var)
        init.zeros_(self.shift)
        init.ones_(self.scale)

    def _axes(self, x):
        raise NotImplementedError

    def forward(self, x):
        if not self._train:
            return (x - self.running_mean) * self.scale / core.sqrt(self.running_var + self.eps) + self.shift
        axes, keep = self._axes(x)
        if F._fused.usable(x, self.scale, self.shift, op='feature_norm'):
            return F._fused.feature_norm(self, x, axes, keep)
        from ... import distributed as dist
        sync = dist.sync_stats_enabled()  # data parallel: statistics of the GLOBAL batch (equal shards)
        mean = x.mean(axes, keepdims=keep)
        if sync:
            mean = dist.dist_mean(mean)
        centered = x - mean
        var = core.mean(core.square(centered), axes, keepdims=keep)
        if sync:
            var = dist.dist_mean(var)
        out = centered / core.sqrt(var + self.eps)
        self.running_mean *= (1 - self.momentum)
        self.running_mean += self.momentum * mean
        self.running_var *= (1 - self.momentum)
        self.running_var += self.momentum * var
        return out * self.scale + self.shift


class BatchNorm1d(_FeatureStatNorm):

    def __init__(self, num_features: int, eps: float = 1e-6, momentum: float = 0.1, device=None, dtype=None) -> None:
        super().__init__()
        self.num_features = num_features
        self._setup(num_features, eps, momentum, device, dtype)

    def _axes(self, x):
        return 0, False

    def __repr__(self) -> str:
        return "{}(num_features={}, momentum={})".format(self.__class__.__name__, self.num_features, self.momentum)


class BatchNorm2d(_FeatureStatNorm):

    def __init__(self, num_features: int, eps: float = 1e-6, momentum: float = 0.1, device=None, dtype=None) -> None:
        super().__init__()
        self.num_features = num_features
        self._setup((1, num_features, 1, 1), eps, momentum, device, dtype)

    def _axes(self, x):
        return (0, 2, 3), True

    def __repr__(self) -> str:
        return "{}(num_features={}, momentum={})".format(self.__class__.__name__, self.num_features, self.momentum)


class LayerNorm(_FeatureStatNorm):

    def __init__(self, normalized_shape, eps: float = 1e-6, momentum: float = 0.1, device=None, dtype=None) -> None:
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape, )
        self.normalized_shape = tuple(normalized_shape)
        self._setup(self.normalized_shape, eps, momentum, device, dtype)

    def _axes(self, x):
        return tuple(range(x.ndim - len(self.normalized_shape))), False


class RMSNorm(Module):
    """x / sqrt(mean(x^2 over the trailing normalized axes) + eps) * weight (reference norm.py:221-248)."""

    def __init__(self, normalized_shape, eps: float = 1e-6, device=None, dtype=None):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape, )
        self.normalized_shape = tuple(normalized_shape)
        self.sum_axis = tuple(-(i + 1) for i in range(len(self.normalized_shape)))
        self.eps = eps
        self.weight = Parameter(empty(self.normalized_shape, device=Device(device), dtype=dtype))
        self.reset_parameters()

    def reset_parameters(self):
        init.ones_(self.weight)

    def forward(self, x):
        if len(self.normalized_shape) == 1 and F._fused.usable(x, self.weight, op='rmsnorm'):
            return F._fused.rmsnorm(x, self.weight, self.eps)
        z = core.square(x).mean(self.sum_axis, keepdims=True)
        return x / core.sqrt(z + self.eps) * self.weight
