from .optimizer import Optimizer, SGD, Adagrad, Adadelta, Adam
from .lr_scheduler import _LRScheduler, ExponentialLR, StepLR, MultiStepLR, CosineAnnealingLR
from . import lr_scheduler
