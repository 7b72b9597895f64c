"""Import path of the reference (pydynet/nn/modules/dropout.py); the classes live in layers.py."""
from .layers import Dropout  # noqa: F401
