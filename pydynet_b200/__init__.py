"""pydynet_b200 — PyDyNet's Tensor / autograd / nn surface on a hand-written sm_100a backend."""
from .cuda import Device
from . import cuda, autograd
from .autograd import no_grad, enable_grad, set_grad_enabled, is_grad_enable
