"""In-place initialisers (reference nn/init.py:6-92). Draws come from ``tensor.xp.random`` — for cuda tensors that is
the host ``np.random`` stream followed by an H2D copy, so a seeded model gets the same weights as the reference."""
import math

from ..autograd import no_grad
from ..core import Tensor

_GAINS = {"linear": 1, "conv1d": 1, "conv2d": 1, "sigmoid": 1, "tanh": 5 / 3, "relu": math.sqrt(2.)}


def calculate_gain(nonlinearity: str, param: float = None) -> float:
    if nonlinearity == "leaky_relu":
        return math.sqrt(2. / (1 + (param if param is not None else 0.01)**2))
    return _GAINS[nonlinearity]


def _calculate_fan(tensor: Tensor):
    assert tensor.ndim >= 2
    fan_in, fan_out = tensor.shape[:2]
    if tensor.ndim > 2:
        rf = math.prod(tensor.shape[2:])
        fan_in, fan_out = fan_in * rf, fan_out * rf
    return fan_in, fan_out


@no_grad()
def uniform_(tensor: Tensor, a=0., b=1.) -> Tensor:
    with tensor.device:
        tensor.data[...] = tensor.xp.random.uniform(a, b, tensor.shape)
    return tensor


@no_grad()
def normal_(tensor: Tensor, mean=0., std=1.) -> Tensor:
    with tensor.device:
        tensor.data[...] = tensor.xp.random.normal(mean, std, size=tensor.shape)
    return tensor


@no_grad()
def constant_(tensor: Tensor, val: float) -> Tensor:
    with tensor.device:
        tensor.data[...] = val
    return tensor


def ones_(tensor: Tensor) -> Tensor:
    return constant_(tensor, 1.)


def zeros_(tensor: Tensor) -> Tensor:
    return constant_(tensor, 0.)


def xavier_uniform_(tensor: Tensor, gain: float = 1.) -> Tensor:
    fan_in, fan_out = _calculate_fan(tensor)
    bound = gain * math.sqrt(6. / (fan_in + fan_out))
    return uniform_(tensor, -bound, bound)


def xavier_normal_(tensor: Tensor, gain: float = 1.) -> Tensor:
    fan_in, fan_out = _calculate_fan(tensor)
    return normal_(tensor, std=gain * math.sqrt(2 / (fan_in + fan_out)))


def _fan(tensor, mode):
    fan_in, fan_out = _calculate_fan(tensor)
    return {"fan_in": fan_in, "fan_out": fan_out}[mode]


def kaiming_uniform_(tensor: Tensor, a: float = 0., mode='fan_in', nonlinearity='relu') -> Tensor:
    bound = calculate_gain(nonlinearity, a) * math.sqrt(3. / _fan(tensor, mode))
    return uniform_(tensor, -bound, bound)


def kaiming_normal_(tensor: Tensor, a: float = 0., mode='fan_in', nonlinearity='relu'):
    return normal_(tensor, std=calculate_gain(nonlinearity, a) / math.sqrt(_fan(tensor, mode)))
