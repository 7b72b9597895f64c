"""Import path of the reference (pydynet/nn/modules/pool.py); the classes live in layers.py."""
from .layers import MaxPool1d, MaxPool2d, AvgPool1d, AvgPool2d  # noqa: F401
