"""Import path of the reference (pydynet/nn/modules/activation.py); the classes live in layers.py."""
from .layers import Sigmoid, Tanh, ReLU, LeakyReLU, Softmax  # noqa: F401
