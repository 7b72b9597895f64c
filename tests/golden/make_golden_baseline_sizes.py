"""Golden vectors AT BASELINE CONFIG SIZES, produced by the UNMODIFIED reference (NumPy CPU path): one training
step (+ the loss after it) of C2 LeNet at batch 256, C4 Transformer encoder d512 / h8 / ffn 1536 / S128 / B8 / V8192, C5 GRU
in512 / h512 / T256 / B32.  At these sizes every GEMM / convolution / attention of the CUDA path runs on the tcgen05 kernels
(the small fixtures of make_golden.py mostly stay below their eligibility thresholds).

Inputs and initial parameters are NOT stored (C4's embedding alone is 16 MB): both sides rebuild them from the same seeds — the
constructors consume the host ``np.random`` stream in the reference's declaration order — and the fixture pins a fingerprint of
every initial parameter (sum, first entries) so a draw-order mismatch is reported as such.  Large gradients keep ~32 evenly
spaced rows.  Run in the build container only: python tests/golden/make_golden_baseline_sizes.py
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
import pydynet as pdn  # noqa: E402  (the reference)
import pydynet.nn as nn  # noqa: E402
import pydynet.nn.functional as F  # noqa: E402
from pydynet.optim import Adam  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def T(a, rg=False, dtype=None):
    a = np.asarray(a)
    return pdn.Tensor(a, dtype=dtype if dtype is not None else a.dtype, requires_grad=rg)


def thin(v):
    v = np.asarray(v)
    return v[::max(1, v.shape[0] // 32)] if v.size > 50_000 else v


def fingerprint(prefix, named):
    out = {}
    for k, p in named:
        a = np.asarray(p.data, dtype=np.float64).ravel()
        out[f"{prefix}fp.{k}"] = np.concatenate([[a.sum(), np.abs(a).sum()], a[:6]])
    return out


def grads(prefix, named):
    return {f"{prefix}g.{k}": thin(np.array(p.grad, copy=True)) for k, p in named if p.requires_grad}


def _exec_ref(path, names, start=None, end=None):
    src = open(path).read().splitlines()
    ns = {"np": np, "pdn": pdn, "nn": nn, "F": F, "DTYPE": f32, "__name__": "ref_model"}
    exec(compile("\n".join(src[start:end]), path, "exec"), ns)
    return [ns[n] for n in names]


def lenet(d):
    (ConvNet, ) = _exec_ref("/root/reference/examples/pydynet/mnist.py", ["ConvNet"], 81, 98)
    np.random.seed(42)
    net = ConvNet()
    rng = np.random.default_rng(1)
    X, y = rng.random((256, 1, 28, 28)).astype(f32), rng.integers(0, 10, 256)
    named = list(net._parameters.items())
    d.update(fingerprint("c2.", named))
    opt = Adam(net.parameters(), lr=1e-4)
    net.train()
    for s in range(2):
        out = net(T(X))
        loss = F.cross_entropy_loss(out, T(y))
        opt.zero_grad()
        loss.backward()
        if s == 0:
            d["c2.logits0"] = out.data.copy()
            d.update(grads("c2.", named))
        d[f"c2.loss{s}"] = loss.data.copy()
        opt.step()


def encoder(d):
    (Transformer, ) = _exec_ref("/root/reference/examples/pydynet/transformer.py", ["Transformer"], 52, 192)
    rng = np.random.default_rng(2)
    X = rng.integers(1, 8192, (8, 128))
    y = rng.choice([-1, 1], 8).astype(f32)
    # the reference against ITSELF in float64 (same seeds, parameters cast up): this network amplifies rounding (softmax + the
    # batch-statistic "LayerNorm"), so the test's bar per tensor is max(1e-4, 3 x the reference's own fp32-vs-fp64
    # distance), measured against the float64 gradients stored here
    np.random.seed(0)
    net64 = Transformer(512, 1, 8, 3, 0.05, 8192, 128)
    net64.word_embedding.reset_parameters()
    for p in net64._parameters.values():
        p.data = p.data.astype(np.float64)
        if p.requires_grad:
            p.grad = np.zeros(p.data.shape, np.float64)
    net64.train()
    out64 = net64(T(X), None)
    loss64 = pdn.log(1 + pdn.exp(-T(y.astype(np.float64)) * pdn.squeeze(out64))).mean()
    loss64.backward()
    d["c4.out0_f64"] = out64.data.copy()
    d.update({f"c4.g64.{k}": thin(np.array(p.grad, copy=True)) for k, p in net64._parameters.items() if p.requires_grad})
    del net64, out64, loss64
    np.random.seed(0)
    net = Transformer(512, 1, 8, 3, 0.05, 8192, 128)
    net.word_embedding.reset_parameters()
    named = list(net._parameters.items())
    d.update(fingerprint("c4.", named))
    opt = Adam(net.parameters(), lr=5e-4)
    net.train()
    for s in range(2):
        out = net(T(X), None)
        loss = pdn.log(1 + pdn.exp(-T(y) * pdn.squeeze(out))).mean()
        opt.zero_grad()
        loss.backward()
        if s == 0:
            d["c4.out0"] = out.data.copy()
            d.update(grads("c4.", named))
        d[f"c4.loss{s}"] = loss.data.copy()
        opt.step()


def gru(d):
    np.random.seed(0)
    rnn = nn.GRU(512, 512, 1, batch_first=True, dtype=f32)
    head = nn.Linear(512, 1, dtype=f32)
    rng = np.random.default_rng(3)
    X, Y = rng.standard_normal((32, 256, 512)).astype(f32), rng.standard_normal((32, 1)).astype(f32)
    named = [("rnn." + k, p) for k, p in rnn._parameters.items()] + [("out." + k, p) for k, p in head._parameters.items()]
    d.update(fingerprint("c5.", named))
    opt = Adam(list(rnn.parameters()) + list(head.parameters()), lr=0.01)
    for s in range(2):
        _, h = rnn(T(X), None)
        pred = head(h[:, 0, :])
        loss = F.mse_loss(pred, T(Y))
        opt.zero_grad()
        loss.backward()
        if s == 0:
            d["c5.pred0"] = pred.data.copy()
            d.update(grads("c5.", named))
        d[f"c5.loss{s}"] = loss.data.copy()
        opt.step()


if __name__ == "__main__":
    d = {}
    for fn in (lenet, encoder, gru):
        fn(d)
        print(fn.__name__, "done", flush=True)
    path = os.path.join(HERE, "baseline_sizes.npz")
    np.savez_compressed(path, **d)
    print(f"baseline_sizes.npz: {len(d)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")
