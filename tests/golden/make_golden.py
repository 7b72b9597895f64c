"""Generates tests/golden/*.npz by running the UNMODIFIED reference (WeltXing/PyDyNet at /root/reference, NumPy CPU path)
on seeded inputs.  Run in the build container only (`python tests/golden/make_golden.py`); the GPU box has no
/root/reference, so the .npz files are committed and the tests read nothing else.

Every fixture stores inputs, ALL parameters (by the reference's ``_parameters`` names) and the reference's outputs /
gradients / post-optimiser parameters, so the tests do not depend on RNG draw order — except `init_*` entries, which pin
exactly that (seeded constructors must produce the reference's initial weights).
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
from pydynet.optim import Adam, SGD, Adagrad, Adadelta  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def T(a, rg=False, dtype=None):
    a = np.asarray(a)
    return pdn.Tensor(a, dtype=dtype if dtype is not None else a.dtype, requires_grad=rg)


def params_of(m, prefix="p."):
    return {prefix + k: v.data.copy() for k, v in m._parameters.items()}


def grads_of(m, prefix="g."):
    return {prefix + k: np.array(v.grad, copy=True) for k, v in m._parameters.items() if v.requires_grad}


def thin(d):
    """Large tensors (LeNet fc1: 1.2 M entries) are pinned on every 16th row only, to keep the fixture small."""
    return {k: (v[::16] if v.size > 100_000 else v) for k, v in d.items()}


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}.npz: {len(d)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


# ------------------------------------------------------------------------------------------------ functional
def gen_functional():
    rng = np.random.default_rng(11)
    d = {}
    # softmax / log_softmax / losses
    x = rng.standard_normal((5, 7, 33)).astype(f32) * 3
    w = rng.standard_normal((5, 7, 33)).astype(f32)
    for name, fn in (("softmax", lambda t: F.softmax(t, axis=-1)), ("log_softmax", lambda t: F.log_softmax(t, axis=-1, keepdims=True))):
        t = T(x, True)
        out = fn(t)
        (out * T(w)).sum().backward()
        d[f"{name}.x"], d[f"{name}.w"], d[f"{name}.out"], d[f"{name}.gx"] = x, w, out.data, t.grad
    t = T(x, True)
    out = F.softmax(t, axis=1)
    (out * T(w)).sum().backward()
    d["softmax_ax1.out"], d["softmax_ax1.gx"] = out.data, t.grad
    logits = rng.standard_normal((9, 10)).astype(f32) * 4
    tgt = rng.integers(0, 10, 9)
    onehot = np.eye(10, dtype=f32)[tgt]
    d["ce.logits"], d["ce.target"], d["ce.onehot"] = logits, tgt, onehot
    for red in ("mean", "sum"):
        t = T(logits, True)
        loss = F.cross_entropy_loss(t, T(tgt), red)
        loss.backward()
        d[f"ce.int.{red}.loss"], d[f"ce.int.{red}.g"] = loss.data, t.grad
        t = T(logits, True)
        loss = F.cross_entropy_loss(t, T(onehot), red)
        loss.backward()
        d[f"ce.onehot.{red}.loss"], d[f"ce.onehot.{red}.g"] = loss.data, t.grad
    a, b = rng.standard_normal((6, 4)).astype(f32), rng.standard_normal((6, 4)).astype(f32)
    d["mse.a"], d["mse.b"] = a, b
    t = T(a, True)
    loss = F.mse_loss(t, T(b))
    loss.backward()
    d["mse.loss"], d["mse.g"] = loss.data, t.grad
    t = T(a, True)
    loss = F.nll_loss(t, T(b), "sum")
    loss.backward()
    d["nll.loss"], d["nll.g"] = loss.data, t.grad
    # activations
    v = rng.standard_normal((4, 9)).astype(f32) * 3
    v[0, :3] = 0.0
    d["act.x"] = v
    for name, fn in (("relu", F.relu), ("leaky", lambda t: F.leaky_relu(t, 0.1)), ("silu", F.silu), ("sigmoid", F.sigmoid), ("tanh", F.tanh)):
        t = T(v, True)
        out = fn(t)
        (out * out).sum().backward()
        d[f"act.{name}.out"], d[f"act.{name}.g"] = out.data, t.grad
    # conv2d / pooling: (N, C, H, W, O, k, stride, pad)
    for i, (N, C, H, W, O, k, s, p) in enumerate([(2, 3, 8, 8, 4, 3, 1, 1), (3, 1, 9, 7, 5, 3, 2, 0), (2, 4, 6, 6, 3, 1, 1, 0), (1, 2, 10, 10, 6, 5, 2, 2)]):
        xx = rng.standard_normal((N, C, H, W)).astype(f32)
        kk = rng.standard_normal((O, C, k, k)).astype(f32)
        bb = rng.standard_normal((1, O, 1, 1)).astype(f32)
        tx, tk, tb = T(xx, True), T(kk, True), T(bb, True)
        out = F.conv2d(tx, tk, p, s) + tb
        ww = rng.standard_normal(out.shape).astype(f32)
        (out * T(ww)).sum().backward()
        d[f"conv{i}.cfg"] = np.array([N, C, H, W, O, k, s, p])
        d[f"conv{i}.x"], d[f"conv{i}.k"], d[f"conv{i}.b"], d[f"conv{i}.w"] = xx, kk, bb, ww
        d[f"conv{i}.out"], d[f"conv{i}.gx"], d[f"conv{i}.gk"], d[f"conv{i}.gb"] = out.data, tx.grad, tk.grad, tb.grad
    for i, (N, C, H, W, k, s, p) in enumerate([(2, 3, 8, 8, 2, 2, 0), (2, 2, 9, 9, 3, 2, 1), (1, 4, 7, 6, 3, 1, 0)]):
        xx = rng.standard_normal((N, C, H, W)).astype(f32)
        xx[0, 0, :2, :2] = 1.5  # a tied window
        d[f"pool{i}.cfg"], d[f"pool{i}.x"] = np.array([N, C, H, W, k, s, p]), xx
        for mode, fn in (("max", F.max_pool2d), ("avg", F.avg_pool2d)):
            tx = T(xx, True)
            out = fn(tx, k, s, p)
            ww = rng.standard_normal(out.shape).astype(f32)
            (out * T(ww)).sum().backward()
            d[f"pool{i}.{mode}.w"], d[f"pool{i}.{mode}.out"], d[f"pool{i}.{mode}.gx"] = ww, out.data, tx.grad
    # 1-d pooling keeps the reference's (odd) reduction axis
    x1 = rng.standard_normal((2, 3, 10)).astype(f32)
    d["pool1d.x"], d["pool1d.max"], d["pool1d.avg"] = x1, F.max_pool1d(T(x1), 2, 2, 0).data, F.avg_pool1d(T(x1), 3, 1, 1).data
    # embedding with padding_idx and duplicate ids (last-write-wins backward)
    emb = nn.Embedding(12, 6, padding_idx=0, dtype=f32)
    np.random.seed(5)
    emb.reset_parameters()
    ids = np.array([[1, 4, 4, 0], [7, 1, 0, 11]])
    out = emb(T(ids))
    ww = rng.standard_normal(out.shape).astype(f32)
    (out * T(ww)).sum().backward()
    d["emb.weight"], d["emb.ids"], d["emb.w"], d["emb.out"], d["emb.g"] = emb.weight.data.copy(), ids, ww, out.data, emb.weight.grad
    save("functional", d)


# ------------------------------------------------------------------------------------------------ modules
def gen_modules():
    rng = np.random.default_rng(12)
    d = {}
    # seeded constructors: RNG draw order
    np.random.seed(3)
    lin, conv, gru, lstm, rnn = nn.Linear(5, 4, dtype=f32), nn.Conv2d(2, 3, 3, dtype=f32), nn.GRUCell(4, 3, dtype=f32), nn.LSTMCell(4, 3, dtype=f32), nn.RNNCell(4, 3, dtype=f32)
    for nm, m in (("lin", lin), ("conv", conv), ("gru", gru), ("lstm", lstm), ("rnn", rnn)):
        d.update(params_of(m, f"init_{nm}."))
    d["init_default_dtype"] = np.array(str(nn.Linear(2, 2).weight.dtype))
    # norms: two training steps (running stats) then eval
    # (BatchNorm2d and LayerNorm with a multi-axis normalized_shape cannot be constructed in the reference: their
    # constructors raise TypeError at norm.py:122 / :196 — nothing to pin.)
    for nm, mod, shape in (("bn1", nn.BatchNorm1d(6, dtype=f32), (8, 6)), ("ln", nn.LayerNorm(6, dtype=f32), (4, 5, 6))):
        mod.scale.data[...] = rng.standard_normal(mod.scale.shape).astype(f32)
        mod.shift.data[...] = rng.standard_normal(mod.shift.shape).astype(f32)
        d.update(params_of(mod, f"{nm}.p0."))
        mod.train()
        for step in range(2):
            xx = (rng.standard_normal(shape) * 2 + 1).astype(f32)
            ww = rng.standard_normal(shape).astype(f32)
            tx = T(xx, True)
            for p in mod.parameters():
                p.zero_grad()
            out = mod(tx)
            (out * T(ww)).sum().backward()
            d[f"{nm}.s{step}.x"], d[f"{nm}.s{step}.w"], d[f"{nm}.s{step}.out"], d[f"{nm}.s{step}.gx"] = xx, ww, out.data, tx.grad
            d[f"{nm}.s{step}.gscale"], d[f"{nm}.s{step}.gshift"] = mod.scale.grad.copy(), mod.shift.grad.copy()
            d[f"{nm}.s{step}.rm"], d[f"{nm}.s{step}.rv"] = mod.running_mean.data.copy(), mod.running_var.data.copy()
        mod.eval()
        d[f"{nm}.eval.out"] = mod(T(xx)).data
        pdn.autograd.set_grad_enabled(True)
    rms = nn.RMSNorm(6, dtype=f32)
    rms.weight.data[...] = rng.standard_normal(6).astype(f32)
    xx, ww = rng.standard_normal((3, 4, 6)).astype(f32), rng.standard_normal((3, 4, 6)).astype(f32)
    tx = T(xx, True)
    out = rms(tx)
    (out * T(ww)).sum().backward()
    d["rms.weight"], d["rms.x"], d["rms.w"], d["rms.out"], d["rms.gx"], d["rms.gw"] = rms.weight.data.copy(), xx, ww, out.data, tx.grad, rms.weight.grad
    # dropout: host RNG stream
    np.random.seed(9)
    dr = nn.Dropout(0.3)
    xx = rng.standard_normal((4, 5)).astype(f32)
    d["drop.x"], d["drop.out"] = xx, dr(T(xx)).data
    # recurrent stacks
    for nm, cls, kw in (("gru", nn.GRU, dict(num_layers=1)), ("gru2b", nn.GRU, dict(num_layers=2, bidirectional=True, batch_first=True)),
                        ("lstm", nn.LSTM, dict(num_layers=1)), ("lstm2b", nn.LSTM, dict(num_layers=2, bidirectional=True)),
                        ("rnn2", nn.RNN, dict(num_layers=2, nonlinearity="relu"))):
        np.random.seed(21)
        mod = cls(5, 7, dtype=f32, **kw)
        Tn, B = 6, 3
        xx = rng.standard_normal((B, Tn, 5) if kw.get("batch_first") else (Tn, B, 5)).astype(f32)
        tx = T(xx, True)
        out, hn = mod(tx)
        if isinstance(hn, tuple):
            hn, cn = hn
            d[f"{nm}.cn"] = cn.data
            extra = (cn * cn).sum()
        else:
            extra = 0
        ww = rng.standard_normal(out.shape).astype(f32)
        ((out * T(ww)).sum() + (hn * hn).sum() + extra).backward()
        d.update(params_of(mod, f"{nm}.p."))
        d.update(grads_of(mod, f"{nm}.g."))
        d[f"{nm}.x"], d[f"{nm}.w"], d[f"{nm}.out"], d[f"{nm}.hn"], d[f"{nm}.gx"] = xx, ww, out.data, hn.data, tx.grad
    # optimisers: 3 steps on a fixed quadratic
    w0 = rng.standard_normal((4, 3)).astype(f32)
    b0 = rng.standard_normal(3).astype(f32)
    xs = rng.standard_normal((3, 5, 4)).astype(f32)
    d["opt.w0"], d["opt.b0"], d["opt.xs"] = w0, b0, xs
    for nm, mk in (("adam", lambda ps: Adam(ps, lr=1e-2, weight_decay=0.01)), ("sgd", lambda ps: SGD(ps, lr=1e-2, momentum=0.9)),
                   ("adagrad", lambda ps: Adagrad(ps, lr=1e-1)), ("adadelta", lambda ps: Adadelta(ps, lr=1.0))):
        w, b = T(w0.copy(), True), T(b0.copy(), True)
        opt = mk([w, b])
        for s in range(3):
            opt.zero_grad()
            ((T(xs[s]) @ w + b)**2).mean().backward()
            opt.step()
            d[f"opt.{nm}.w{s}"], d[f"opt.{nm}.b{s}"] = w.data.copy(), b.data.copy()
    save("modules", d)


# ------------------------------------------------------------------------------------------------ models
def _exec_ref(path, names, start=None, end=None):
    """Executes a slice of a reference example file (model class definitions only — the scripts' data loading needs
    packages absent here) in a namespace where `pydynet` is the reference."""
    src = open(path).read().splitlines()
    code = "\n".join(src[start:end])
    ns = {"np": np, "pdn": pdn, "nn": nn, "F": F, "DTYPE": f32, "__name__": "ref_model"}
    exec(compile(code, path, "exec"), ns)
    return [ns[n] for n in names]


def gen_lenet():
    (ConvNet, ) = _exec_ref("/root/reference/examples/pydynet/mnist.py", ["ConvNet"], 81, 98)
    np.random.seed(42)
    net = ConvNet()
    rng = np.random.default_rng(1)
    X = rng.random((8, 1, 28, 28)).astype(f32)
    y = rng.integers(0, 10, 8)
    d = {"X": X, "y": y}
    d.update(params_of(net, "p0."))
    opt = Adam(net.parameters(), lr=1e-3)
    net.train()
    for s in range(2):
        out = net(T(X))
        loss = F.cross_entropy_loss(out, T(y))
        opt.zero_grad()
        loss.backward()
        if s == 0:
            d["logits0"] = out.data.copy()
            d.update(thin(grads_of(net, "g0.")))
        d[f"loss{s}"] = loss.data.copy()
        opt.step()
    d.update(thin(params_of(net, "p2.")))
    save("lenet", d)


def gen_transformer():
    SelfAttention, TransformerBlock, spe, construct_mask, Transformer = _exec_ref(
        "/root/reference/examples/pydynet/transformer.py",
        ["SelfAttention", "TransformerBlock", "sinusoidal_positional_encoding", "construct_mask", "Transformer"], 52, 192)
    np.random.seed(0)
    net = Transformer(32, 1, 4, 3, 0.05, 40, 12)
    net.word_embedding.reset_parameters()
    rng = np.random.default_rng(2)
    X = rng.integers(1, 40, (6, 12))
    X[0, 9:] = 0  # padded tail on one row
    X[3, 10:] = 0
    y = rng.choice([-1, 1], 6).astype(f32)
    d = {"X": X, "y": y}
    d.update(params_of(net, "p0."))
    opt = Adam(net.parameters(), lr=5e-4)
    net.train()
    for s in range(2):
        out = net(T(X), construct_mask(T(X)))
        loss = pdn.log(1 + pdn.exp(-T(y) * pdn.squeeze(out))).mean()
        opt.zero_grad()
        loss.backward()
        if s == 0:
            d["out0"] = out.data.copy()
            d.update(grads_of(net, "g0."))
        d[f"loss{s}"] = loss.data.copy()
        opt.step()
    d.update(params_of(net, "p2."))
    net.eval()
    d["eval_out"] = net(T(X), construct_mask(T(X))).data.copy()
    pdn.autograd.set_grad_enabled(True)
    save("transformer", d)


def gen_gru():
    np.random.seed(0)
    gru = nn.GRU(6, 10, 1, batch_first=True, dtype=f32)
    head = nn.Linear(10, 1, dtype=f32)
    rng = np.random.default_rng(3)
    X = rng.standard_normal((5, 20, 6)).astype(f32)
    Y = rng.standard_normal((5, 1)).astype(f32)
    d = {"X": X, "Y": Y}
    d.update(params_of(gru, "p0.rnn."))
    d.update(params_of(head, "p0.out."))
    params = list(gru.parameters()) + list(head.parameters())
    opt = Adam(params, lr=0.01)
    for s in range(2):
        _, h = gru(T(X), None)
        pred = head(h[:, 0, :])
        loss = F.mse_loss(pred, T(Y))
        opt.zero_grad()
        loss.backward()
        if s == 0:
            d["pred0"] = pred.data.copy()
            d.update(grads_of(gru, "g0.rnn."))
            d.update(grads_of(head, "g0.out."))
        d[f"loss{s}"] = loss.data.copy()
        opt.step()
    d.update(params_of(gru, "p2.rnn."))
    d.update(params_of(head, "p2.out."))
    save("gru", d)


def gen_llama():
    sys.path.insert(0, "/root/reference")
    from llm.llama.model import Llama
    np.random.seed(0)
    V, D, H, FF, S, B, L = 96, 48, 4, 128, 64, 2, 2
    net = Llama(V, D, H, FF, S, B, L, f32)
    for name, p in net._parameters.items():
        if "cache" in name or "freqs" in name:
            continue
        if "norm" in name:
            p.data[...] = (1 + 0.1 * np.random.randn(*p.shape)).astype(f32)
        else:
            p.data[...] = (0.08 * np.random.randn(*p.shape)).astype(f32)
    d = {"cfg": np.array([V, D, H, FF, S, B, L])}
    d.update({k: v for k, v in params_of(net, "p.").items() if "cache" not in k})
    # fine-tune style forward/backward over all positions (training mode: no KV cache)
    ids = np.array([[1, 5, 9, 33, 2, 70], [4, 4, 80, 95, 0, 17]])
    tgt = np.roll(ids, -1, axis=1)
    net.train(True)
    logits = net.forward_logits(T(ids))
    loss = nn.CrossEntropyLoss()(logits.reshape(-1, V), T(tgt.reshape(-1)))
    loss.backward()
    d["ft.ids"], d["ft.tgt"], d["ft.logits"], d["ft.loss"] = ids, tgt, logits.data.copy(), loss.data.copy()
    for k in ("lm_head.weight", "layers.1.ffn.down.weight", "layers.0.attention.Q.weight", "layers.0.input_norm.weight", "tok_embedding.weight"):
        d["ft.g." + k] = np.array(net._parameters[k].grad, copy=True)
    # greedy generation with the KV cache (eval mode), batch of 2 prompts
    net.eval()
    prompt = np.array([[1, 7, 20, 3], [9, 9, 41, 60]])
    with pdn.no_grad():
        d["gen.prefill_logits"] = net(T(prompt), 0).data.copy()
        for layer in net.layers:
            layer.attention.cache_k.data[...] = 0
            layer.attention.cache_v.data[...] = 0
        toks, margins = [], []
        for nid in net.generate(T(prompt), 40):
            toks.append(nid.data.copy())
        # top-1/top-2 margins of the oracle, teacher-forced replay, to know which steps are well conditioned
    d["gen.prompt"], d["gen.tokens"] = prompt, np.concatenate(toks, axis=1)
    d["gen.cache_k0"] = net.layers[0].attention.cache_k.data.copy()
    pdn.autograd.set_grad_enabled(True)
    save("llama", d)


if __name__ == "__main__":
    gen_functional()
    gen_modules()
    gen_lenet()
    gen_transformer()
    gen_gru()
    gen_llama()
