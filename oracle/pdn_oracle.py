"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain-NumPy restatement of the reference's (WeltXing/PyDyNet, NumPy path) algorithms on the dense-tensor hot path,
written as explicit forward / backward formulas (no autograd tape), each citing the reference file:line it follows.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm may import it — as the checker or
as the timed CPU baseline — never the product: pydynet_b200's cuda path does not import this module and fails loudly if
libpdn_b200.so is missing.

Parity pin: every function here is checked in tests/test_oracle.py against tests/golden/*.npz, which were produced by
running the UNMODIFIED reference in the build container (tests/golden/make_golden.py).  The reference is pure Python, so
there is nothing to compile into oracle/_ref; the reference's arithmetic bottoms out in NumPy >= 2.0 + its bundled
OpenBLAS (requirements.txt:1), which is also what this file calls.

All functions take and return NumPy arrays and preserve the input dtype (fp32 in, fp32 arithmetic) unless noted.
"""
import numpy as np


# ---------------------------------------------------------------------------------- elementwise pieces ----------
def sigmoid(x):
    """Piecewise overflow-safe logistic (reference pydynet/core/tensor.py:996-1002)."""
    out = np.empty_like(x)
    pos = x > 0
    out[pos] = 1 / (1 + np.exp(-x[pos]))
    out[~pos] = 1 - 1 / (1 + np.exp(x[~pos]))
    return out


def tanh(x):
    """Piecewise tanh (reference tensor.py:1009-1015)."""
    out = np.empty_like(x)
    pos = x > 0
    out[pos] = 2 / (1 + np.exp(-2 * x[pos])) - 1
    out[~pos] = 1 - 2 / (1 + np.exp(2 * x[~pos]))
    return out


def relu(x):
    """maximum(0, x); relu'(0) = 1 because ties send the gradient to both operands (tensor.py:808-815, functional.py:31-32)."""
    return np.maximum(x.dtype.type(0), x)


def relu_grad(x, g):
    return (np.maximum(x.dtype.type(0), x) == x) * g


def silu(x):
    """x / (1 + exp(-x)) (functional.py:39-40)."""
    return x / (1 + np.exp(-x))


# ---------------------------------------------------------------------------------- matmul ----------------------
def matmul_fwd_bwd(a, b, g):
    """out = a @ b ; da = g @ bᵀ ; db = aᵀ @ g summed over broadcast batch dims (tensor.py:657-676 + engine un-broadcast
    tensor.py:360-370). >= 2-D operands."""
    out = a @ b
    da = g @ np.swapaxes(b, -1, -2)
    db = np.swapaxes(a, -1, -2) @ g
    return out, _unbroadcast(da, a.shape), _unbroadcast(db, b.shape)


def _unbroadcast(g, shape):
    extra = g.ndim - len(shape)
    if extra:
        g = g.sum(axis=tuple(range(extra)))
    axes = tuple(i for i, s in enumerate(shape) if s == 1 and g.shape[i] != 1)
    if axes:
        g = g.sum(axis=axes, keepdims=True)
    return g


# ---------------------------------------------------------------------------------- softmax / losses ------------
def softmax(x, axis=-1):
    """max (no grad) -> sub -> exp -> sum -> div (functional.py:43-49)."""
    e = np.exp(x - x.max(axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def softmax_bwd(y, g, axis=-1):
    return y * (g - (g * y).sum(axis=axis, keepdims=True))


def log_softmax(x, axis=-1):
    """functional.py:52-58 with keepdims=True."""
    s = x - x.max(axis, keepdims=True)
    return s - np.log(np.exp(s).sum(axis=axis, keepdims=True))


def log_softmax_bwd(y, g, axis=-1):
    return g - np.exp(y) * g.sum(axis=axis, keepdims=True)


def cross_entropy(logits, target, reduction="mean"):
    """functional.py:364-381: shift by the GLOBAL max, log-sum-exp over axis 1; integer targets pick one entry per row,
    one-hot targets weight all N*C entries (so 'mean' divides by N*C). Returns (loss, dlogits)."""
    shifted = logits - logits.max()
    lse = np.log(np.exp(shifted).sum(1, keepdims=True))
    nls = lse - shifted
    p = np.exp(-nls)
    if target.ndim == 1:
        n = logits.shape[0]
        picked = nls[np.arange(n), target]
        scale = 1.0 / n if reduction == "mean" else 1.0
        onehot = np.zeros_like(logits)
        onehot[np.arange(n), target] = 1
        return (picked.mean() if reduction == "mean" else picked.sum()).astype(logits.dtype), ((p - onehot) * scale).astype(logits.dtype)
    w = nls * target
    scale = 1.0 / w.size if reduction == "mean" else 1.0
    grad = (p * target.sum(1, keepdims=True) - target) * scale
    return (w.mean() if reduction == "mean" else w.sum()).astype(logits.dtype), grad.astype(logits.dtype)


def mse(a, b):
    """mean((a-b)^2) and its gradient wrt a (functional.py:342-350)."""
    d = a - b
    return (d * d).mean(), 2 * d / d.size


# ---------------------------------------------------------------------------------- conv / pool -----------------
def _windows(xp, k, stride):
    """Strided sliding-window view (N, C, k, k, oh, ow) of a padded NCHW array (functional.py:211-222)."""
    N, C, H, W = xp.shape
    oh, ow = (H - k) // stride + 1, (W - k) // stride + 1
    s0, s1, s2, s3 = xp.strides
    return np.lib.stride_tricks.as_strided(xp, (N, C, k, k, oh, ow), (s0, s1, s2, s3, s2 * stride, s3 * stride)), oh, ow


def _pad(x, p):
    return np.pad(x, [(0, 0), (0, 0), (p, p), (p, p)], "constant") if p else x


def _col2im(gcol6, xshape_padded, k, stride, p):
    """Scatter-add of window gradients back onto the padded input, then crop (functional.py:224-232, 247-251)."""
    gx = np.zeros(xshape_padded, dtype=gcol6.dtype)
    view, _, _ = _windows(gx, k, stride)
    np.add.at(view, (Ellipsis, ), gcol6)
    return gx[:, :, p:gx.shape[2] - p, p:gx.shape[3] - p] if p else gx


def conv2d_fwd_bwd(x, w, bias, g, stride=1, pad=0):
    """im2col convolution (functional.py:254-281) + (1,O,1,1) bias (conv.py:99-103). Returns out, dx, dw, dbias;
    pass g=None for forward only."""
    N = x.shape[0]
    O, C, k, _ = w.shape
    xp = _pad(x, pad)
    win, oh, ow = _windows(xp, k, stride)
    col = win.transpose(0, 4, 5, 1, 2, 3).reshape(N * oh * ow, -1)
    wmat = w.reshape(O, -1).T
    out = (col @ wmat).reshape(N, oh, ow, O).transpose(0, 3, 1, 2)
    if bias is not None:
        out = out + bias
    if g is None:
        return out
    g2 = g.transpose(0, 2, 3, 1).reshape(N * oh * ow, O)
    dw = (col.T @ g2).T.reshape(w.shape)
    dcol = (g2 @ wmat.T).reshape(N, oh, ow, C, k, k).transpose(0, 3, 4, 5, 1, 2)
    dx = _col2im(dcol, xp.shape, k, stride, pad)
    db = g.sum(axis=(0, 2, 3), keepdims=True) if bias is not None else None
    return out, dx, dw, db


def pool2d_fwd_bwd(x, k, stride, pad, mode, g=None):
    """max / avg pooling through the same im2col (functional.py:284-339); zero padding takes part in max/mean; max
    backward gives the FULL gradient to every element equal to the window max (tensor.py:741-747)."""
    N, C = x.shape[:2]
    xp = _pad(x, pad)
    win, oh, ow = _windows(xp, k, stride)
    col = win.transpose(0, 4, 5, 1, 2, 3).reshape(-1, k * k)
    red = col.max(1) if mode == "max" else col.mean(1)
    out = red.reshape(N, oh, ow, C).transpose(0, 3, 1, 2)
    if g is None:
        return out
    gflat = g.transpose(0, 2, 3, 1).reshape(-1, 1)
    gcol = (col == red[:, None]) * gflat if mode == "max" else np.broadcast_to(gflat / (k * k), col.shape)
    gcol6 = gcol.reshape(N, oh, ow, C, k, k).transpose(0, 3, 4, 5, 1, 2)
    return out, _col2im(np.ascontiguousarray(gcol6), xp.shape, k, stride, pad)


# ---------------------------------------------------------------------------------- norms -----------------------
def feature_norm_fwd_bwd(x, scale, shift, axes, eps=1e-6, g=None):
    """Shared body of BatchNorm1d/2d and the reference's "LayerNorm" in training mode (norm.py:58-73, 132-147, 203-218):
    per-feature mean / biased variance over `axes` (for LayerNorm: the LEADING axes), y = (x-mean)/sqrt(var+eps)*scale+shift.
    Returns (y, mean, var) or, with g, (y, mean, var, dx, dscale, dshift)."""
    keep = x.ndim == scale.ndim
    mean = x.mean(axes, keepdims=keep)
    c = x - mean
    var = (c * c).mean(axes, keepdims=keep)
    rstd = 1 / np.sqrt(var + eps)
    xhat = c * rstd
    y = xhat * scale + shift
    if g is None:
        return y, mean, var
    red = axes if isinstance(axes, tuple) else (axes, )
    dshift = g.sum(red, keepdims=keep)
    dscale = (g * xhat).sum(red, keepdims=keep)
    m = x.size // mean.size
    dx = scale * rstd * (g - dshift / m - xhat * dscale / m)
    return y, mean, var, dx, dscale, dshift


def rmsnorm_fwd_bwd(x, w, eps=1e-6, g=None):
    """x / sqrt(mean(x^2, last axis) + eps) * w (norm.py:245-248)."""
    ms = (x * x).mean(-1, keepdims=True)
    r = 1 / np.sqrt(ms + eps)
    y = x * r * w
    if g is None:
        return y
    gw = g * w
    dx = r * gw - x * r**3 * (gw * x).mean(-1, keepdims=True)
    dw = (g * x * r).reshape(-1, x.shape[-1]).sum(0)
    return y, dx, dw


# ---------------------------------------------------------------------------------- recurrent -------------------
def gru_seq_fwd_bwd(x, h0, Wx1, Wh1, Wx2, Wh2, b1, b2, g_out=None, g_hn=None):
    """GRU over time (cell rnn.py:529-544, loop :702-708). x [T,B,I], h0 [B,H]. zr = sigmoid(x Wx1 + h Wh1 + b1), z = first
    half, r = second half; n = tanh(x Wx2 + (r*h) Wh2 + b2); h' = (1-z) h + z n. Returns hs [T,B,H] or, with gradients of
    the outputs / final state, (hs, dx, dh0, dWx1, dWh1, dWx2, dWh2, db1, db2) via explicit BPTT."""
    T, B, _ = x.shape
    H = h0.shape[1]
    hs, zs, rs, ns = [], [], [], []
    h = h0
    for t in range(T):
        zr = sigmoid(x[t] @ Wx1 + h @ Wh1 + b1)
        z, r = zr[:, :H], zr[:, H:]
        n = tanh(x[t] @ Wx2 + (r * h) @ Wh2 + b2)
        h = (1 - z) * h + z * n
        hs.append(h); zs.append(z); rs.append(r); ns.append(n)
    hs = np.stack(hs)
    if g_out is None and g_hn is None:
        return hs
    dx = np.zeros_like(x)
    dWx1, dWh1, dWx2, dWh2 = (np.zeros_like(w) for w in (Wx1, Wh1, Wx2, Wh2))
    db1, db2 = np.zeros_like(b1), np.zeros_like(b2)
    dh = np.zeros_like(h0) if g_hn is None else g_hn.copy()
    for t in reversed(range(T)):
        if g_out is not None:
            dh = dh + g_out[t]
        hp = hs[t - 1] if t > 0 else h0
        z, r, n = zs[t], rs[t], ns[t]
        dn = dh * z
        dz = dh * (n - hp)
        dhp = dh * (1 - z)
        dl2 = dn * (1 - n * n)
        dWx2 += x[t].T @ dl2
        dWh2 += (r * hp).T @ dl2
        db2 += dl2.sum(0)
        drh = dl2 @ Wh2.T
        dr = drh * hp
        dhp = dhp + drh * r
        dl1 = np.concatenate([dz * z * (1 - z), dr * r * (1 - r)], axis=1)
        dWx1 += x[t].T @ dl1
        dWh1 += hp.T @ dl1
        db1 += dl1.sum(0)
        dx[t] = dl1 @ Wx1.T + dl2 @ Wx2.T
        dh = dhp + dl1 @ Wh1.T
    return hs, dx, dh, dWx1, dWh1, dWx2, dWh2, db1, db2


def lstm_seq_fwd(x, h0, c0, Wx, Wh, b):
    """LSTM over time (rnn.py:268-288): lin = x Wx + h Wh + b; sigmoid(first 3H) -> f, i, o; tanh(last H) -> g;
    c' = f c + i g; h' = o tanh(c'). Returns hs, cs."""
    H = h0.shape[1]
    h, c, hs, cs = h0, c0, [], []
    for t in range(x.shape[0]):
        lin = x[t] @ Wx + h @ Wh + b
        fio = sigmoid(lin[:, :3 * H])
        f, i, o = fio[:, :H], fio[:, H:2 * H], fio[:, 2 * H:]
        c = f * c + i * tanh(lin[:, 3 * H:])
        h = o * tanh(c)
        hs.append(h); cs.append(c)
    return np.stack(hs), np.stack(cs)


# ---------------------------------------------------------------------------------- Adam ------------------------
def adam_step(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """One Adam update, in place (optimizer.py:185-196); t is the shared step counter starting at 1."""
    g = g + wd * p
    m *= b1
    m += (1 - b1) * g
    v *= b2
    v += (1 - b2) * g**2
    a_t = np.sqrt(1 - b2**t) / (1 - b1**t)
    p -= (lr * a_t * m / (v**0.5 + eps)).astype(p.dtype)


# ---------------------------------------------------------------------------------- attention -------------------
def attention(q, k, v, mask=None, scale=1.0):
    """softmax(q kᵀ * scale + mask) v for [B,H,L,D] operands (llm/llama/model.py:112-118)."""
    s = q @ np.swapaxes(k, -1, -2) * q.dtype.type(scale)
    if mask is not None:
        s = s + mask
    return softmax(s, -1) @ v


# ---------------------------------------------------------------------------------- Llama -----------------------
def rope_tables(head_dim, max_seq_len, dtype=np.float32, base=10000):
    """cos/sin [max_seq_len, head_dim/2] (llm/llama/model.py:10-20)."""
    inv = 1.0 / (base**(np.arange(0, head_dim, 2)[:head_dim // 2] / head_dim))
    f = np.outer(np.arange(max_seq_len), inv).astype(dtype)
    return np.cos(f), np.sin(f)


def rope(x, cos, sin):
    """Interleaved-pair rotation of [B,L,H,D] by per-position angles [L,D/2] (model.py:23-44)."""
    xr, xi = x[..., 0::2], x[..., 1::2]
    c, s = cos[:, None, :], sin[:, None, :]
    out = np.empty_like(x)
    out[..., 0::2] = xr * c - xi * s
    out[..., 1::2] = xr * s + xi * c
    return out


class LlamaOracle:
    """Functional Llama (model.py:47-269) over a dict of weights keyed by the reference's ``_parameters`` names, with the
    per-layer KV cache and the reference's greedy ``generate`` bookkeeping (decode step i runs the token of position L+i-1
    with start_pos = L+i, model.py:258-267)."""

    def __init__(self, params, n_heads, max_seq_len, max_batch, n_layers, eps=1e-6):
        self.p = params
        self.H, self.S, self.L = n_heads, max_seq_len, n_layers
        D = params["tok_embedding.weight"].shape[1]
        self.dim, self.hd = D, D // n_heads
        dt = params["tok_embedding.weight"].dtype
        self.cos, self.sin = rope_tables(self.hd, max_seq_len, dt)
        self.ck = [np.zeros((max_batch, max_seq_len, n_heads, self.hd), dt) for _ in range(n_layers)]
        self.cv = [np.zeros((max_batch, max_seq_len, n_heads, self.hd), dt) for _ in range(n_layers)]
        self.eps = eps

    def hidden(self, ids, start_pos, use_cache=True):
        p = self.p
        B, L = ids.shape
        h = p["tok_embedding.weight"][ids]
        cos, sin = self.cos[start_pos:start_pos + L], self.sin[start_pos:start_pos + L]
        mask = None
        if L > 1:
            mask = np.concatenate([np.zeros((L, start_pos)), np.triu(np.full((L, L), -np.inf), k=1)], axis=1).astype(h.dtype)
        for i in range(self.L):
            pre = f"layers.{i}."
            nx = rmsnorm_fwd_bwd(h, p[pre + "input_norm.weight"], self.eps)
            q = (nx @ p[pre + "attention.Q.weight"]).reshape(B, L, self.H, self.hd)
            k = (nx @ p[pre + "attention.K.weight"]).reshape(B, L, self.H, self.hd)
            v = (nx @ p[pre + "attention.V.weight"]).reshape(B, L, self.H, self.hd)
            q, k = rope(q, cos, sin), rope(k, cos, sin)
            if use_cache:
                self.ck[i][:B, start_pos:start_pos + L] = k
                self.cv[i][:B, start_pos:start_pos + L] = v
                k, v = self.ck[i][:B, :start_pos + L], self.cv[i][:B, :start_pos + L]
            o = attention(q.transpose(0, 2, 1, 3), k.transpose(0, 2, 1, 3), v.transpose(0, 2, 1, 3), mask, 1.0 / np.sqrt(self.hd))
            z = h + o.transpose(0, 2, 1, 3).reshape(B, L, -1) @ p[pre + "attention.O.weight"]
            nz = rmsnorm_fwd_bwd(z, p[pre + "post_attn_norm.weight"], self.eps)
            ff = (silu(nz @ p[pre + "ffn.gate.weight"]) * (nz @ p[pre + "ffn.up.weight"])) @ p[pre + "ffn.down.weight"]
            h = z + ff
        return rmsnorm_fwd_bwd(h, p["norm.weight"], self.eps)

    def logits_all(self, ids):
        """Training-mode forward over all positions, no cache (model.py:209-211)."""
        return self.hidden(ids, 0, use_cache=False) @ self.p["lm_head.weight"] + self.p["lm_head.bias"]

    def step(self, ids, start_pos):
        """model.py:254-256: logits of the last position only, [B,1,V]."""
        return self.hidden(ids, start_pos)[:, [-1], :] @ self.p["lm_head.weight"] + self.p["lm_head.bias"]

    def generate(self, prompt, max_total_len):
        """Greedy ids, [B, max_total_len - L] (model.py:258-269)."""
        _, L = prompt.shape
        out, nxt = [], None
        for i, cur in enumerate(range(L, max_total_len)):
            logits = self.step(prompt, 0) if i == 0 else self.step(nxt, cur)
            nxt = logits[:, -1, :].argmax(-1, keepdims=True)
            out.append(nxt)
        return np.concatenate(out, axis=1)


    def generate_with_margins(self, prompt, max_total_len):
        """generate() that also returns, per step and sequence, the relative top-1/top-2 logit margin
        (l1 - l2) / max|logit| — the quantity the parity rule of SURVEY.md §8(c) is stated on."""
        _, L = prompt.shape
        out, mar, nxt = [], [], None
        for i, cur in enumerate(range(L, max_total_len)):
            logits = (self.step(prompt, 0) if i == 0 else self.step(nxt, cur))[:, -1, :]
            nxt = logits.argmax(-1, keepdims=True)
            top2 = np.partition(logits, -2, axis=-1)[:, -2:]
            mar.append((top2[:, 1] - top2[:, 0]) / np.maximum(np.abs(logits).max(axis=-1), 1e-30))
            out.append(nxt)
        return np.concatenate(out, axis=1), np.stack(mar, axis=1)


def check_greedy_tokens(got, ref_tokens, ref_margins, tol=1e-4):
    """The token parity rule: every sequence must equal the oracle's greedy ids EXACTLY up to (excluding) the first
    step at which the oracle's own top-1/top-2 margin is below ``tol`` (an fp32 near-tie: a different but equally accurate summation
    order may legitimately pick the other token, after which the contexts differ). Returns (n_exact_sequences, n_diverged_at_near_tie);
    raises AssertionError on any mismatch at a step whose margin is above ``tol``."""
    got, ref_tokens = np.asarray(got), np.asarray(ref_tokens)
    assert got.shape == ref_tokens.shape, (got.shape, ref_tokens.shape)
    exact = near = 0
    for b in range(got.shape[0]):
        bad = np.nonzero(got[b] != ref_tokens[b])[0]
        if bad.size == 0:
            exact += 1
            continue
        t = int(bad[0])
        assert ref_margins[b, t] < tol, (f"sequence {b} differs from the oracle at step {t} (got {got[b, t]}, oracle {ref_tokens[b, t]}) "
                                         f"although the oracle's top-1/top-2 margin there is {ref_margins[b, t]:.3e} >= {tol:g}")
        near += 1
    return exact, near


def synthetic_llama_params(V, D, H, FF, n_layers, seed=0, std=0.05, dtype=np.float32):
    """Random-init weights of the BASELINE config-3 architecture (no checkpoints are reachable offline): every matrix
    N(0, std), norm weights 1, lm_head bias N(0, std) — SURVEY.md §8(d) C3."""
    rng = np.random.default_rng(seed)
    n = lambda *s: (rng.standard_normal(s) * std).astype(dtype)
    p = {"tok_embedding.weight": n(V, D), "norm.weight": np.ones(D, dtype), "lm_head.weight": n(D, V), "lm_head.bias": n(V)}
    for i in range(n_layers):
        pre = f"layers.{i}."
        for nm in "QKVO":
            p[pre + f"attention.{nm}.weight"] = n(D, D)
        p[pre + "ffn.up.weight"], p[pre + "ffn.gate.weight"], p[pre + "ffn.down.weight"] = n(D, FF), n(D, FF), n(FF, D)
        p[pre + "input_norm.weight"] = np.ones(D, dtype)
        p[pre + "post_attn_norm.weight"] = np.ones(D, dtype)
    return p


# ---------------------------------------------------------------------------------- LeNet -----------------------
def lenet_loss_and_grads(p, X, y):
    """Forward + backward of the BASELINE config-2 ConvNet (examples/pydynet/mnist.py:82-98) with CE loss; p maps the
    reference parameter names to arrays. Returns (logits, loss, grads dict)."""
    a1 = conv2d_fwd_bwd(X, p["conv1.weight"], p["conv1.bias"], None, 1, 1)
    r1 = relu(a1)
    p1 = pool2d_fwd_bwd(r1, 2, 2, 0, "max")
    a2 = conv2d_fwd_bwd(p1, p["conv2.weight"], p["conv2.bias"], None, 1, 1)
    r2 = relu(a2)
    p2 = pool2d_fwd_bwd(r2, 2, 2, 0, "max")
    f = p2.reshape(-1, 2450)
    h1 = f @ p["fc1.weight"] + p["fc1.bias"]
    rh = relu(h1)
    logits = rh @ p["fc2.weight"] + p["fc2.bias"]
    loss, dlog = cross_entropy(logits, y)
    g = {"fc2.weight": rh.T @ dlog, "fc2.bias": dlog.sum(0)}
    drh = dlog @ p["fc2.weight"].T
    dh1 = relu_grad(h1, drh)
    g["fc1.weight"], g["fc1.bias"] = f.T @ dh1, dh1.sum(0)
    dp2 = (dh1 @ p["fc1.weight"].T).reshape(p2.shape)
    _, dr2 = pool2d_fwd_bwd(r2, 2, 2, 0, "max", dp2)
    da2 = relu_grad(a2, dr2)
    _, dp1, g["conv2.weight"], g["conv2.bias"] = conv2d_fwd_bwd(p1, p["conv2.weight"], p["conv2.bias"], da2, 1, 1)
    _, dr1 = pool2d_fwd_bwd(r1, 2, 2, 0, "max", dp1)
    da1 = relu_grad(a1, dr1)
    _, _, g["conv1.weight"], g["conv1.bias"] = conv2d_fwd_bwd(X, p["conv1.weight"], p["conv1.bias"], da1, 1, 1)
    return logits, loss, g
