"""LSTM and plain RNN at sizes the persistent whole-sequence kernels take (k_lstm_persist_fwd / _bwd, k_rnn_persist<0|1>, H % 64 == 0)
against what the
unmodified reference produced (tests/golden/lstm_sizes.npz, tests/golden/make_golden_lstm.py): forward outputs, h_n, c_n and the
gradients wrt x, h0, c0 and every parameter, with given initial states and gradients flowing into the whole output sequence, h_n and
c_n.  Case a: 1 layer in192 / h256 / T48 / B40 (ragged batch tile); case b: 2 layers bidirectional h128 / T20 / B72 (two batch
tiles); case r: RNN tanh 1 layer h128 / T24 / B40; case s: RNN relu 2 layers bidirectional h64 / T16 / B72.  Tolerance 1e-4 normwise (BF16x3 recurrent products, fp32 gates)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydynet_b200 as pdn  # noqa: E402
import pydynet_b200.nn as nn  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from lstm_cases import CASES, inputs, thin  # noqa: E402  (shared with the fixture's generating script)

G = np.load(os.path.join(ROOT, "tests", "golden", "lstm_sizes.npz"))
DEVICES = ["cpu", pytest.param("cuda:0", marks=pytest.mark.gpu)]
f32 = np.float32
LOOP = os.environ.get("PDN_GRU_PERSIST") == "0"  # the per-step host loop was asked for: same results, no persistent launches


def close(got, ref, what, rtol=1e-4):
    got = got.numpy() if isinstance(got, pdn.Tensor) else (got.get() if hasattr(got, "get") else got)
    got, ref = thin(np.asarray(got, np.float64)), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    assert err < rtol, f"{what}: normwise rel err {err:.3e} >= {rtol:g}"


def T(a, dev, rg=False):
    return pdn.Tensor(np.asarray(a), dtype=f32, requires_grad=rg, device=dev)


@pytest.mark.parametrize("dev", DEVICES)
@pytest.mark.parametrize("nm", list(CASES))
def test_lstm_sequence_against_reference(dev, nm):
    c = CASES[nm]
    np.random.seed(11)
    mod = getattr(nn, c["cls"])(c["I"], c["H"], dtype=f32, **c["kw"])
    lstm = c["cls"] == "LSTM"
    kern = "lstm_persist" if lstm else "rnn_persist"
    mod.to(dev)
    params = list(mod.parameters())
    for i, p in enumerate(params):
        a = p.numpy().astype(np.float64).ravel()
        np.testing.assert_allclose(np.concatenate([[a.sum(), np.abs(a).sum()], a[:6]]), G[f"{nm}.fp.{i}"], rtol=1e-6, atol=1e-7,
                                   err_msg=f"initial parameter {i}: RNG draw order differs from the reference")
    x, h0, c0, w = inputs(c, 5)
    tx, th, tc = T(x, dev, True), T(h0, dev, True), T(c0, dev, True)
    if dev != "cpu":
        from pydynet_b200.backend import lib
        lib.watch_launches(kern + "_fwd")
    if lstm:
        out, (hn, cn) = mod(tx, (th, tc))
        loss = (out * T(w, dev)).sum() + (hn * hn).sum() + (cn * cn).sum() * 0.5
    else:
        out, hn = mod(tx, th)
        loss = (out * T(w, dev)).sum() + (hn * hn).sum()
    if dev != "cpu":
        nd = c["kw"].get("num_layers", 1) * (2 if c["kw"].get("bidirectional") else 1)
        assert lib.watched_launch_count() == (0 if LOOP else nd), "the persistent forward kernel did not run"
        lib.watch_launches(kern + "_bwd")
    loss.backward()
    if dev != "cpu":
        assert lib.watched_launch_count() == (0 if LOOP else nd), "the persistent backward kernel did not run"
        lib.watch_launches(None)
    close(out, G[f"{nm}.out"], "out")
    close(hn, G[f"{nm}.hn"], "hn")
    np.testing.assert_allclose(float(loss.item()), float(G[f"{nm}.loss"]), rtol=1e-4)
    close(tx.grad, G[f"{nm}.dx"], "dx")
    close(th.grad, G[f"{nm}.dh0"], "dh0")
    if lstm:
        close(cn, G[f"{nm}.cn"], "cn")
        close(tc.grad, G[f"{nm}.dc0"], "dc0")
    for i, p in enumerate(params):
        close(p.grad, G[f"{nm}.g.{i}"], f"grad of parameter {i}")
