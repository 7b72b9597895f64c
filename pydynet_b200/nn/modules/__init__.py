from .layers import (Sigmoid, Tanh, ReLU, LeakyReLU, Softmax, Conv1d, Conv2d, MaxPool1d, MaxPool2d, AvgPool1d, AvgPool2d, Dropout,
                     Linear, Embedding, MSELoss, NLLLoss, CrossEntropyLoss)
from .norm import BatchNorm1d, BatchNorm2d, LayerNorm, RMSNorm
from .module import Module, Sequential, ModuleList
from .rnn import RNN, LSTM, GRU, RNNCell, LSTMCell, GRUCell
from . import activation, conv, dropout, linear, loss, pool  # the reference's submodule import paths

__all__ = [
    "Sigmoid", "Tanh", "ReLU", "LeakyReLU", "Softmax", "BatchNorm1d", "BatchNorm2d", "LayerNorm", "RMSNorm", "Conv1d", "Conv2d",
    "MaxPool1d", "MaxPool2d", "AvgPool1d", "AvgPool2d", "Dropout", "Linear", "Embedding", "MSELoss", "NLLLoss",
    "CrossEntropyLoss", "Module", "Sequential", "ModuleList", "RNN", "LSTM", "GRU", "RNNCell", "LSTMCell", "GRUCell"
]
