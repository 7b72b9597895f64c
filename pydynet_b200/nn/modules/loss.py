"""Import path of the reference (pydynet/nn/modules/loss.py); the classes live in layers.py."""
from .layers import Loss, MSELoss, NLLLoss, CrossEntropyLoss  # noqa: F401
