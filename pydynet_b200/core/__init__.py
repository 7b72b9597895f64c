from .tensor import (Tensor, add, sub, mul, div, pow, matmul, abs, sum, mean, min, max, argmax, argmin, maximum, minimum,
                     exp, log, sign, reshape, transpose, swapaxes, concat, sigmoid, tanh, _get_slice, _UnaryOperator,
                     _BinaryOperator)
from .function import sqrt, square, vsplit, hsplit, dsplit, split, unsqueeze, squeeze
