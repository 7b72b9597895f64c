"""Recurrent layers (reference nn/modules/rnn.py:13-723): RNNCell/LSTMCell/GRUCell and their sequence wrappers with
multi-layer / bidirectional / batch_first handling.

Cell maths (the parity contract):
  RNN   h' = act(x Wx + h Wh + b)
  LSTM  lin = x Wx + h Wh + b ; first 3H columns -> sigmoid -> f, i, o (that order) ; last H -> tanh -> g ;
        c' = f*c + i*g ; h' = o*tanh(c')                                               (reference rnn.py:280-288)
  GRU   zr = sigmoid(x Wx1 + h Wh1 + b1) ; z, r = halves (z first) ; n = tanh(x Wx2 + (r*h) Wh2 + b2) ;
        h' = (1-z)*h + z*n                                                             (reference rnn.py:537-544)

The reference has three near-identical sequence classes; here one ``_Recurrent`` base drives any cell type.  On a cuda
device an fp32 GRU/LSTM layer runs as ONE fused sequence node (hoisted input GEMM over all T + persistent recurrent
kernel, nn/_fused.py) instead of ~18 tape entries per time step.
"""
import math

from .module import Module
from .. import init
from .. import functional as F
from ..parameter import Parameter
from ... import core
from ...cuda import Device
from ...special import empty, zeros


class _Cell(Module):
    _gates = 1  # multiple of hidden_size produced by the fused projections

    def __init__(self, input_size: int, hidden_size: int, bias: bool = True, device=None, dtype=None) -> None:
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.kwargs = {"device": Device(device), "dtype": dtype}
        self.has_bias = bias

    def _check_state(self, x, s, what="hidden"):
        ok = (x.ndim == 1 and s.shape == (self.hidden_size, )) or (x.ndim == 2 and s.shape == (x.shape[0], self.hidden_size))
        assert ok, "Wrong {} state input!".format(what)

    def init_hidden(self, x):
        assert x.ndim in {1, 2}
        shape = self.hidden_size if x.ndim == 1 else (x.shape[0], self.hidden_size)
        return zeros(shape, **self.kwargs)

    def move(self, device):
        self.kwargs['device'] = device
        return super().move(device)

    def __repr__(self) -> str:
        return "{}({}, {}, bias={})".format(self.__class__.__name__, self.input_size, self.hidden_size, self.has_bias)


class RNNCell(_Cell):

    def __init__(self, input_size, hidden_size, bias=True, nonlinearity='tanh', device=None, dtype=None) -> None:
        super().__init__(input_size, hidden_size, bias, device, dtype)
        self.nonlinearity = nonlinearity
        self.fn = {'tanh': F.tanh, 'relu': F.relu}[nonlinearity]
        self.Wx = Parameter(empty((input_size, hidden_size), **self.kwargs))
        self.Wh = Parameter(empty((hidden_size, hidden_size), **self.kwargs))
        if bias:
            self.bias = Parameter(empty(hidden_size, **self.kwargs))
        self.reset_paramters()

    def reset_paramters(self):
        bound = math.sqrt(1 / self.hidden_size)
        for p in (self.Wx, self.Wh) + ((self.bias, ) if self.has_bias else ()):
            init.uniform_(p, -bound, bound)

    reset_parameters = reset_paramters

    def forward(self, x, h=None):
        if h is None:
            h = self.init_hidden(x)
        else:
            self._check_state(x, h)
        lin = x @ self.Wx + h @ self.Wh
        if self.has_bias:
            lin = lin + self.bias
        return self.fn(lin)

    def __repr__(self) -> str:
        return "{}({}, {}, bias={}, nonlinearity={})".format(self.__class__.__name__, self.input_size, self.hidden_size,
                                                             self.has_bias, self.nonlinearity)


class LSTMCell(_Cell):

    def __init__(self, input_size, hidden_size, bias=True, device=None, dtype=None) -> None:
        super().__init__(input_size, hidden_size, bias, device, dtype)
        self.Wx = Parameter(empty((input_size, 4 * hidden_size), **self.kwargs))
        self.Wh = Parameter(empty((hidden_size, 4 * hidden_size), **self.kwargs))
        if bias:
            self.bias = Parameter(empty(4 * hidden_size, **self.kwargs))
        self.reset_paramters()

    def reset_paramters(self):
        bound = math.sqrt(1 / self.hidden_size)
        for p in (self.Wx, self.Wh) + ((self.bias, ) if self.has_bias else ()):
            init.uniform_(p, -bound, bound)

    reset_parameters = reset_paramters

    def forward(self, x, hx=None):
        if hx is None:
            h, c = self.init_hidden(x), self.init_hidden(x)
        else:
            h, c = hx
            self._check_state(x, h)
            self._check_state(x, c, "cell")
        H = self.hidden_size
        lin = x @ self.Wx + h @ self.Wh
        if self.has_bias:
            lin = lin + self.bias
        fio, g = core.split(lin, [3 * H], axis=-1)
        f, i, o = core.split(F.sigmoid(fio), 3, axis=-1)
        c = f * c + i * F.tanh(g)
        return o * F.tanh(c), c


class GRUCell(_Cell):

    def __init__(self, input_size, hidden_size, bias=True, device=None, dtype=None) -> None:
        super().__init__(input_size, hidden_size, bias, device, dtype)
        H = hidden_size
        self.Wx1 = Parameter(empty((input_size, 2 * H), **self.kwargs))
        self.Wh1 = Parameter(empty((H, 2 * H), **self.kwargs))
        self.Wx2 = Parameter(empty((input_size, H), **self.kwargs))
        self.Wh2 = Parameter(empty((H, H), **self.kwargs))
        if bias:
            self.bias1 = Parameter(empty(2 * H, **self.kwargs))
            self.bias2 = Parameter(empty(H, **self.kwargs))
        self.reset_parameters()

    def reset_parameters(self):
        bound = math.sqrt(1 / self.hidden_size)
        order = (self.Wx1, self.Wx2, self.Wh1, self.Wh2) + ((self.bias1, self.bias2) if self.has_bias else ())
        for p in order:  # draw order of the reference (rnn.py:546-554)
            init.uniform_(p, -bound, bound)

    def forward(self, x, h=None):
        if h is None:
            h = self.init_hidden(x)
        else:
            self._check_state(x, h)
        lin1 = x @ self.Wx1 + h @ self.Wh1
        if self.has_bias:
            lin1 = lin1 + self.bias1
        z, r = core.split(F.sigmoid(lin1), 2, axis=-1)
        lin2 = x @ self.Wx2 + (r * h) @ self.Wh2
        if self.has_bias:
            lin2 = lin2 + self.bias2
        return (1 - z) * h + z * F.tanh(lin2)


class _Recurrent(Module):
    """Sequence driver shared by RNN / LSTM / GRU."""
    _prefix = "rnn"
    _has_cell_state = False

    def _make_cell(self, in_size):
        raise NotImplementedError

    def __init__(self, input_size, hidden_size, num_layers=1, bias=True, batch_first=False, bidirectional=False, device=None,
                 dtype=None) -> None:
        super().__init__()
        assert num_layers > 0
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.has_bias, self.batch_first, self.bidirectional = bias, batch_first, bidirectional
        self.kwargs = {"device": Device(device), "dtype": dtype}
        sizes = [input_size] + [hidden_size] * (num_layers - 1)
        self._fwd_cells, self._bwd_cells = [], []
        for i in range(num_layers):
            cell = self._make_cell(sizes[i])
            setattr(self, '{}_{}'.format(self._prefix, i), cell)
            self._fwd_cells.append(cell)
        if bidirectional:
            for i in range(num_layers):
                cell = self._make_cell(sizes[i])
                setattr(self, 'r{}_{}'.format(self._prefix, i), cell)
                self._bwd_cells.append(cell)

    def reset_parameters(self):
        for cell in self._fwd_cells + self._bwd_cells:
            cell.reset_parameters()

    def init_hidden(self, x):
        assert x.ndim in {2, 3}
        d = 2 if self.bidirectional else 1
        shape = (d * self.num_layers, self.hidden_size) if x.ndim == 2 else (d * self.num_layers, x.shape[1], self.hidden_size)
        return zeros(shape, **self.kwargs)

    def move(self, device):
        self.kwargs['device'] = device
        return super().move(device)

    # one layer, one direction: returns (states over time [T, ...] per state kind, last state per kind [1, ...])
    def _run(self, cell, x, state):
        fused = F._fused.sequence(cell, x, state)
        if fused is not None:
            return fused
        T = x.shape[0]
        tracks = [[] for _ in state]
        cur = state if self._has_cell_state else state[0]
        for t in range(T):
            cur = cell(x[t], cur)
            for k, s in enumerate(cur if self._has_cell_state else (cur, )):
                tracks[k].append(core.unsqueeze(s, axis=0))
        return [core.concat(tr) for tr in tracks], [tr[-1] for tr in tracks]

    def forward(self, x, hx=None):
        swap = self.batch_first and x.ndim == 3
        if swap:
            x = x.swapaxes(0, 1)
        nstate = 2 if self._has_cell_state else 1
        if hx is None:
            states = [self.init_hidden(x) for _ in range(nstate)]
        else:
            states = list(hx) if self._has_cell_state else [hx]
            d = 2 if self.bidirectional else 1
            want = (d * self.num_layers, self.hidden_size) if x.ndim == 2 else (d * self.num_layers, x.shape[1], self.hidden_size)
            for s, what in zip(states, ("hidden", "cell")):
                assert s.shape == want, "Wrong {} state input!".format(what)
        L = self.num_layers
        f_in, b_in = x, (x[::-1] if self.bidirectional else None)
        f_last, b_last = [], []
        for i in range(L):
            f_seq, last = self._run(self._fwd_cells[i], f_in, [s[i] for s in states])
            f_in = f_seq[0]
            f_last.append(last)
            if self.bidirectional:
                # each direction stacks on its own outputs; the reverse stack stays in reversed time order between
                # layers (reference rnn.py:151-166)
                b_seq, last = self._run(self._bwd_cells[i], b_in, [s[i + L] for s in states])
                b_in = b_seq[0]
                b_last.append(last)
        if self.bidirectional:
            output = core.concat([f_in, b_in[::-1]], axis=-1)
        else:
            output = f_in
        finals = []
        for k in range(nstate):
            parts = [l[k] for l in f_last] + [l[k] for l in b_last]
            finals.append(parts[0] if len(parts) == 1 else core.concat(parts))
        if swap:
            output = output.swapaxes(0, 1)
            finals = [f.swapaxes(0, 1) for f in finals]
        return (output, tuple(finals)) if self._has_cell_state else (output, finals[0])

    def __repr__(self) -> str:
        return "{}({}, {}, num_layers={}, bias={}, batch_first={}, bidirectional={})".format(
            self.__class__.__name__, self.input_size, self.hidden_size, self.num_layers, self.has_bias, self.batch_first,
            self.bidirectional)


class RNN(_Recurrent):
    _prefix = "rnn"

    def __init__(self, input_size, hidden_size, num_layers=1, nonlinearity='tanh', bias=True, batch_first=False,
                 bidirectional=False, device=None, dtype=None) -> None:
        self.nonlinearity = nonlinearity
        super().__init__(input_size, hidden_size, num_layers, bias, batch_first, bidirectional, device, dtype)
        self.RNNCells, self.rRNNCells = self._fwd_cells, self._bwd_cells

    def _make_cell(self, in_size):
        return RNNCell(in_size, self.hidden_size, self.has_bias, self.nonlinearity, **self.kwargs)


class LSTM(_Recurrent):
    _prefix = "lstm"
    _has_cell_state = True

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.LSTMCells, self.rLSTMCells = self._fwd_cells, self._bwd_cells

    def _make_cell(self, in_size):
        return LSTMCell(in_size, self.hidden_size, self.has_bias, **self.kwargs)


class GRU(_Recurrent):
    _prefix = "gru"

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.GRUCells, self.rGRUCells = self._fwd_cells, self._bwd_cells

    def _make_cell(self, in_size):
        return GRUCell(in_size, self.hidden_size, self.has_bias, **self.kwargs)
