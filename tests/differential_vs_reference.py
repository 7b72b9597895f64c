"""Live differential check against the UNMODIFIED reference on the cpu device (run by tests/test_differential.py where /root/reference is
mounted): the same seeded module is built in both packages, driven for two optimisation steps, and outputs / input gradients /
updated parameters must agree to 1e-5 — module families and optimizer variants beyond the committed golden fixtures (1-D conv / pools,
LeakyReLU, Softmax, Dropout's RNG stream, padded Embedding, bidirectional multi-layer LSTM, batch-first GRU, SGD momentum / nesterov /
weight decay, Adagrad, Adadelta)."""
import sys, warnings, importlib, traceback
warnings.filterwarnings("ignore")
import numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "/root/reference")
import pydynet as R
import pydynet.nn as Rnn, pydynet.nn.functional as RF, pydynet.optim as Ropt
import pydynet_b200 as O
import pydynet_b200.nn as Onn, pydynet_b200.nn.functional as OF, pydynet_b200.optim as Oopt

def run(pkg, nn, F, opt, build, x_np, steps=2, optname="Adam", optkw=None, train=True, seed=0):
    np.random.seed(seed)
    net = build(nn)
    net.train() if train else net.eval()
    pkg.autograd.set_grad_enabled(True) if hasattr(pkg, "autograd") else None
    params = [p for p in net.parameters()] if hasattr(net, "parameters") else []
    o = getattr(opt, optname)(params, **(optkw or {})) if params and any(p.requires_grad for p in params) else None
    outs = []
    for s in range(steps):
        xs = [pkg.Tensor(a, dtype=a.dtype, requires_grad=(a.dtype.kind == 'f')) for a in x_np]
        np.random.seed(100 + s)
        out = net(*xs)
        if isinstance(out, tuple): out = out[0]
        loss = (out * out).sum() if out.dtype.kind == 'f' else out.sum()
        if o: o.zero_grad()
        loss.backward()
        outs.append(out.numpy()); outs.append(xs[0].grad.copy() if xs[0].requires_grad and xs[0].grad is not None else np.zeros(1))
        if o: o.step()
    for n, p in getattr(net, "_parameters", {}).items():
        outs.append(np.array(p.data, copy=True))
    return outs

CASES = {
 "LeakyReLU": (lambda nn: nn.LeakyReLU(0.2), [np.random.RandomState(1).randn(4, 5).astype(np.float32)]),
 "Softmax": (lambda nn: nn.Softmax(-1), [np.random.RandomState(2).randn(3, 6).astype(np.float32)]),
 "Conv1d": (lambda nn: nn.Conv1d(3, 4, 3, 1, 1, dtype=np.float32), [np.random.RandomState(3).randn(2, 3, 3).astype(np.float32)]),
 "MaxPool1d": (lambda nn: nn.MaxPool1d(2, 2, 0), [np.random.RandomState(4).randn(2, 3, 8).astype(np.float32)]),
 "AvgPool1d": (lambda nn: nn.AvgPool1d(2, 2, 0), [np.random.RandomState(5).randn(2, 3, 8).astype(np.float32)]),
 "AvgPool2d": (lambda nn: nn.AvgPool2d(2, 2, 1), [np.random.RandomState(6).randn(2, 3, 6, 6).astype(np.float32)]),
 "Dropout": (lambda nn: nn.Dropout(0.3), [np.random.RandomState(7).randn(5, 7).astype(np.float32)]),
 "BatchNorm1d": (lambda nn: nn.BatchNorm1d(6, dtype=np.float32), [np.random.RandomState(8).randn(9, 6).astype(np.float32)]),
 "RMSNorm": (lambda nn: nn.RMSNorm(6, dtype=np.float32), [np.random.RandomState(9).randn(4, 3, 6).astype(np.float32)]),
 "Embedding_pad": (lambda nn: (lambda e: (e.reset_parameters(), e)[1])(nn.Embedding(10, 4, padding_idx=0, dtype=np.float32)), [np.array([[1, 0, 3], [0, 9, 9]])]),  # (the constructor leaves the table uninitialised in the reference)
 "LSTM_bi2": (lambda nn: nn.LSTM(3, 4, 2, bidirectional=True, dtype=np.float32), [np.random.RandomState(10).randn(5, 2, 3).astype(np.float32)]),
 "GRU_bf": (lambda nn: nn.GRU(3, 4, 1, batch_first=True, dtype=np.float32), [np.random.RandomState(11).randn(2, 5, 3).astype(np.float32)]),
 "RNN_2": (lambda nn: nn.RNN(3, 4, 2, dtype=np.float32), [np.random.RandomState(12).randn(5, 2, 3).astype(np.float32)]),
 "Seq": (lambda nn: nn.Sequential(nn.Linear(5, 7, dtype=np.float32), nn.Tanh(), nn.Linear(7, 2, dtype=np.float32), nn.Sigmoid()), [np.random.RandomState(13).randn(6, 5).astype(np.float32)]),
}
OPTS = [("SGD", dict(lr=0.1)), ("SGD", dict(lr=0.1, momentum=0.9)), ("SGD", dict(lr=0.1, momentum=0.9, nesterov=True, weight_decay=0.01)),
        ("Adagrad", dict(lr=0.1, weight_decay=0.01)), ("Adadelta", dict(weight_decay=0.01)), ("Adam", dict(lr=0.01, weight_decay=0.01))]
bad = 0
for name, (build, xs) in CASES.items():
    for optname, kw in (OPTS if name == "Seq" else [("Adam", dict(lr=0.01))]):
        try:
            a = run(R, Rnn, RF, Ropt, build, xs, optname=optname, optkw=kw)
        except Exception as e:
            print(f"{name}/{optname}: reference raises {type(e).__name__}: {str(e)[:80]}")
            try:
                run(O, Onn, OF, Oopt, build, xs, optname=optname, optkw=kw); print("   ours: runs")
            except Exception as e2:
                print(f"   ours raises {type(e2).__name__}: {str(e2)[:80]}")
            continue
        try:
            b = run(O, Onn, OF, Oopt, build, xs, optname=optname, optkw=kw)
        except Exception as e:
            print(f"{name}/{optname}: OURS raises {type(e).__name__}: {str(e)[:120]}"); bad += 1; continue
        ok = len(a) == len(b) and all(x.shape == y.shape and np.allclose(x, y, rtol=1e-5, atol=1e-6, equal_nan=True) for x, y in zip(a, b))
        if not ok:
            bad += 1
            for i, (x, y) in enumerate(zip(a, b)):
                if x.shape != y.shape or not np.allclose(x, y, rtol=1e-5, atol=1e-6, equal_nan=True):
                    print(f"{name}/{optname} {kw}: MISMATCH item {i} shapes {x.shape} {y.shape} maxdiff {np.abs(x - y).max() if x.shape == y.shape else 'n/a'}")
                    break
        else:
            print(f"{name}/{optname}: ok ({len(a)} arrays)")
print("bad:", bad)
sys.exit(1 if bad else 0)
