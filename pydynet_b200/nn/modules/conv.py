"""Import path of the reference (pydynet/nn/modules/conv.py); the classes live in layers.py."""
from .layers import Conv1d, Conv2d  # noqa: F401
