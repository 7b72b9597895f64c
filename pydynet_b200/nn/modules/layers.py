"""Parameterised and stateless layers: Linear, Embedding, Conv1d/2d, pooling, Dropout, activations, losses
(reference nn/modules/linear.py:12-80, conv.py:10-114, pool.py:5-82, dropout.py:6-21, activation.py:6-77,
loss.py:6-33).  Initialisers are drawn in the reference's declaration order so seeded models match bit for bit."""
import math

from .module import Module
from ..parameter import Parameter
from .. import init, functional as F
from ...autograd import no_grad
from ...cuda import Device
from ...special import empty, rand


class Linear(Module):
    """y = x @ W + b with W stored (in_features, out_features)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        kw = {"device": Device(device), "dtype": dtype}
        self.weight = Parameter(empty((in_features, out_features), **kw))
        self.bias = Parameter(empty(out_features, **kw)) if bias else None
        self.reset_paramters()

    def reset_paramters(self):  # (sic) the reference spells it this way
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan(self.weight)
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            init.uniform_(self.bias, -bound, bound)

    reset_parameters = reset_paramters

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)

    def __repr__(self) -> str:
        return "Linear(in_features={}, out_features={}, bias={})".format(self.in_features, self.out_features, self.bias is not None)


class Embedding(Module):
    """Row gather; NOT initialised by the constructor (reference linear.py:49-67) — call reset_parameters()."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx=None, device=None, dtype=None) -> None:
        super().__init__()
        self.num_embedding, self.embedding_dim, self.padding_idx = num_embeddings, embedding_dim, padding_idx
        self.weight = Parameter(empty((num_embeddings, embedding_dim), device=Device(device), dtype=dtype))

    def forward(self, x):
        return F.embedding(x, self.weight, self.padding_idx)

    def reset_parameters(self) -> None:
        init.normal_(self.weight)
        # the reference's "_fill_padding_idx_with_zero" assigns to the .data of a temporary slice tensor and therefore
        # leaves the weight row untouched (linear.py:73-79); kept as a no-op so seeded weights match.

    def __repr__(self) -> str:
        return "Embedding({}, {}, padding_idx={})".format(self.num_embedding, self.embedding_dim, self.padding_idx)


class _ConvNd(Module):
    _nsp = 2

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, device=None, dtype=None):
        super().__init__()
        kw = {"device": Device(device), "dtype": dtype}
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.padding, self.stride = kernel_size, padding, stride
        self.weight = Parameter(empty((out_channels, in_channels) + (kernel_size, ) * self._nsp, **kw))
        self.bias = Parameter(empty((1, out_channels) + (1, ) * self._nsp, **kw)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan(self.weight)
            if fan_in != 0:
                bound = 1 / math.sqrt(fan_in)
                init.uniform_(self.bias, -bound, bound)

    def __repr__(self) -> str:
        return "{}(in_channels={}, out_channels={}, kernel_size={}, padding={}, stride={}, bias={})".format(
            self.__class__.__name__, self.in_channels, self.out_channels, self.kernel_size, self.padding, self.stride,
            self.bias is not None)


class Conv1d(_ConvNd):
    _nsp = 1

    def forward(self, x):
        out = F.conv1d(x, self.weight, self.padding, self.stride)
        return out + self.bias if self.bias is not None else out


class Conv2d(_ConvNd):
    _nsp = 2

    def forward(self, x):
        if self.bias is not None and F._fused.usable(x, self.weight, self.bias, op='conv2d'):
            return F._fused.conv2d(x, self.weight, self.padding, self.stride, self.bias)
        out = F.conv2d(x, self.weight, self.padding, self.stride)
        return out + self.bias if self.bias is not None else out


class _Pool(Module):
    _fn = None

    def __init__(self, kernel_size: int, stride: int, padding: int) -> None:
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def forward(self, x):
        return type(self)._fn(x, self.kernel_size, self.stride, self.padding)

    def __repr__(self) -> str:
        return "{}(kernel_size={}, stride={}, padding={})".format(self.__class__.__name__, self.kernel_size, self.stride, self.padding)


class MaxPool1d(_Pool):
    _fn = staticmethod(F.max_pool1d)


class AvgPool1d(_Pool):
    _fn = staticmethod(F.avg_pool1d)


class MaxPool2d(_Pool):
    _fn = staticmethod(F.max_pool2d)


class AvgPool2d(_Pool):
    _fn = staticmethod(F.avg_pool2d)


class Dropout(Module):
    """Inverted dropout; the mask is drawn on the host (np.random.rand, the reference's stream) and moved over."""

    def __init__(self, p: float = 0.5) -> None:
        super().__init__()
        assert p >= 0 and p < 1
        self.p = p

    def forward(self, x):
        if self._train:
            mask = rand(*x.shape, device=x.device) < 1 - self.p
            return x * mask.astype(x.dtype) / (1 - self.p)
        return x

    def __repr__(self) -> str:
        return "{}(p={})".format(self.__class__.__name__, self.p)


class _Stateless(Module):

    def __repr__(self) -> str:
        return "{}()".format(self.__class__.__name__)


class Sigmoid(_Stateless):

    def forward(self, x):
        return F.sigmoid(x)


class Tanh(_Stateless):

    def forward(self, x):
        return F.tanh(x)


class ReLU(_Stateless):

    def forward(self, x):
        return F.relu(x)


class LeakyReLU(Module):

    def __init__(self, alpha: float = 0.1) -> None:
        super().__init__()
        self.alpha = float(alpha)

    def forward(self, x):
        return F.leaky_relu(x, self.alpha)

    def __repr__(self) -> str:
        return "{}(alpha={})".format(self.__class__.__name__, self.alpha)


class Softmax(Module):

    def __init__(self, axis=None) -> None:
        super().__init__()
        self.axis = axis

    def forward(self, x):
        return F.softmax(x, self.axis)

    def __repr__(self) -> str:
        return "{}(axis={})".format(self.__class__.__name__, self.axis)


class Loss(Module):

    def __init__(self, reduction='mean') -> None:
        super().__init__()
        assert reduction in {'mean', 'sum'}
        self.reduction = reduction


class MSELoss(Loss):

    def forward(self, y_pred, y_true):
        return F.mse_loss(y_pred, y_true, reduction=self.reduction)


class NLLLoss(Loss):

    def forward(self, y_pred, y_true):
        return F.nll_loss(y_pred, y_true, reduction=self.reduction)


class CrossEntropyLoss(Loss):

    def forward(self, y_pred, y_true):
        return F.cross_entropy_loss(y_pred, y_true, reduction=self.reduction)
