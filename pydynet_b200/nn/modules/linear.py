"""Import path of the reference (pydynet/nn/modules/linear.py); the classes live in layers.py."""
from .layers import Linear, Embedding  # noqa: F401
