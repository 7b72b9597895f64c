from .modules import *
from .parameter import Parameter
from . import init
from . import functional
