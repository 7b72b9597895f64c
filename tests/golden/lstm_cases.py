"""Shapes, seeded inputs and fixture thinning shared by make_golden_lstm.py (the reference side) and tests/test_lstm_sizes.py
(LSTM and plain-RNN cases at sizes the persistent whole-sequence kernels take)."""
import numpy as np

f32 = np.float32
CASES = {"a": dict(cls="LSTM", I=192, H=256, T=48, B=40, kw=dict(num_layers=1)),
         "b": dict(cls="LSTM", I=64, H=128, T=20, B=72, kw=dict(num_layers=2, bidirectional=True)),
         "r": dict(cls="RNN", I=96, H=128, T=24, B=40, kw=dict(num_layers=1, nonlinearity="tanh")),
         "s": dict(cls="RNN", I=32, H=64, T=16, B=72, kw=dict(num_layers=2, bidirectional=True, nonlinearity="relu"))}


def inputs(c, seed):
    rng = np.random.default_rng(seed)
    nd = c["kw"].get("num_layers", 1) * (2 if c["kw"].get("bidirectional") else 1)
    x = rng.standard_normal((c["T"], c["B"], c["I"])).astype(f32)
    h0 = (0.5 * rng.standard_normal((nd, c["B"], c["H"]))).astype(f32)
    c0 = (0.5 * rng.standard_normal((nd, c["B"], c["H"]))).astype(f32)
    w = rng.standard_normal((c["T"], c["B"], c["H"] * (2 if c["kw"].get("bidirectional") else 1))).astype(f32)
    return x, h0, c0, w


def thin(v):
    """large arrays keep ~8 evenly spaced slices of their first axis (time steps / weight rows), 3-D ones also every 3rd batch row"""
    v = np.asarray(v)
    if v.size > 20_000:
        v = v[::max(1, v.shape[0] // 8)]
        if v.ndim == 3 and v.size > 20_000:
            v = v[:, ::3]
    return v


