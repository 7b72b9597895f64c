"""Golden vectors for the LSTM (and the plain RNN: cases r, s — tanh 1 layer H128 / T24 / B40, relu 2 layers bidirectional H64 / T16 /
B72) at a size the persistent whole-sequence kernels take (H % 64 == 0): one forward + backward of the
UNMODIFIED reference (NumPy CPU path) for LSTM in192 / h256 / T48 / B40 (B not a multiple of the 64-row batch tile) with given
(h0, c0), a loss that feeds gradients into the whole output sequence, h_n and c_n, and the same for a 2-layer bidirectional LSTM
h128 / T20 / B72 (two batch tiles).  Inputs / initial parameters are rebuilt from seeds on both sides; fingerprints of the initial
parameters are pinned; large arrays are thinned along their first axis (``thin``).  Run in the build container only: python tests/golden/make_golden_lstm.py
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pydynet as pdn  # noqa: E402  (the reference)
import pydynet.nn as nn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
from lstm_cases import CASES, inputs, thin  # noqa: E402


def T(a, rg=False):
    a = np.asarray(a)
    return pdn.Tensor(a, dtype=a.dtype, requires_grad=rg)


if __name__ == "__main__":
    d = {}
    for nm, c in CASES.items():
        np.random.seed(11)
        mod = getattr(nn, c["cls"])(c["I"], c["H"], dtype=f32, **c["kw"])
        x, h0, c0, w = inputs(c, 5)
        params = list(mod.parameters())
        for i, p in enumerate(params):
            a = np.asarray(p.data, np.float64).ravel()
            d[f"{nm}.fp.{i}"] = np.concatenate([[a.sum(), np.abs(a).sum()], a[:6]])
        tx, th, tc = T(x, True), T(h0, True), T(c0, True)
        if c["cls"] == "LSTM":
            out, (hn, cn) = mod(tx, (th, tc))
            loss = (out * T(w)).sum() + (hn * hn).sum() + (cn * cn).sum() * 0.5
        else:
            out, hn = mod(tx, th)
            loss = (out * T(w)).sum() + (hn * hn).sum()
        loss.backward()
        d[f"{nm}.out"], d[f"{nm}.hn"], d[f"{nm}.loss"] = out.data.copy(), hn.data.copy(), loss.data.copy()
        d[f"{nm}.dx"], d[f"{nm}.dh0"] = tx.grad.copy(), th.grad.copy()
        if c["cls"] == "LSTM":
            d[f"{nm}.cn"], d[f"{nm}.dc0"] = cn.data.copy(), tc.grad.copy()
        for i, p in enumerate(params):
            d[f"{nm}.g.{i}"] = np.array(p.grad, copy=True)
        print(nm, "done", float(loss.data), flush=True)
    d = {k: thin(v) for k, v in d.items()}
    path = os.path.join(HERE, "lstm_sizes.npz")
    np.savez_compressed(path, **d)
    print(f"lstm_sizes.npz: {len(d)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")
