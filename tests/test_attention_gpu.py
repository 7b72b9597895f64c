"""GPU parity of the fused attention operators (tensor-core flash kernels and the warp-per-query kernel) against a NumPy
fp64 evaluation of the reference's score -> softmax -> PV chain (llm/llama/model.py:112-118) and its analytic gradients.
Tolerance 1e-4 normwise (BASELINE north_star) for outputs and all three input gradients; masks with -inf entries included."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(q, k, v, mask, scale, g):
    """[B,L,H,D] inputs in float64; returns out [B,Lq,H*D] and grads."""
    Q, K, V = (t.transpose(0, 2, 1, 3) for t in (q, k, v))
    s = Q @ K.transpose(0, 1, 3, 2) * scale
    if mask is not None:
        s = s + mask
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(-1, keepdims=True)
    o = p @ V
    B, H, Lq, D = o.shape
    out = o.transpose(0, 2, 1, 3).reshape(B, Lq, H * D)
    go = g.reshape(B, Lq, H, D).transpose(0, 2, 1, 3)
    dV = p.transpose(0, 1, 3, 2) @ go
    dP = go @ V.transpose(0, 1, 3, 2)
    dS = p * (dP - (dP * p).sum(-1, keepdims=True)) * scale
    dQ, dK = dS @ K, dS.transpose(0, 1, 3, 2) @ Q
    return out, dQ.transpose(0, 2, 1, 3), dK.transpose(0, 2, 1, 3), dV.transpose(0, 2, 1, 3)


def _err(got, ref):
    return float(np.linalg.norm(got.astype(np.float64) - ref) / max(np.linalg.norm(ref), 1e-30))


CASES = [  # B, H, Lq, Lk, D, mask kind
    (2, 3, 200, 150, 48, "causal"), (2, 4, 256, 256, 64, None), (3, 2, 130, 70, 64, "pad"), (1, 2, 64, 300, 32, None), (2, 2, 96, 96, 16, "causal")
]


@pytest.mark.parametrize("impl", ["tc", "ffma"])
@pytest.mark.parametrize("case", CASES)
def test_attention_fwd_bwd(impl, case):
    import pydynet_b200 as pdn
    from pydynet_b200.nn import _fused
    B, H, Lq, Lk, D, mk = case
    rng = np.random.default_rng(hash(case) % 1000)
    q, k, v = (rng.standard_normal((B, L, H, D)).astype(np.float32) for L in (Lq, Lk, Lk))
    g = rng.standard_normal((B, Lq, H * D)).astype(np.float32)
    mask = None
    if mk == "causal":
        mask = np.triu(np.full((Lq, Lk), -np.inf, np.float32), k=1 + max(0, Lk - Lq))
    elif mk == "pad":
        mask = np.zeros((B, 1, 1, Lk), np.float32)
        mask[0, ..., Lk - 17:] = -np.inf
        mask[2, ..., Lk - 3:] = -np.inf
    scale = 1.0 / np.sqrt(D)
    ref = _ref(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), None if mask is None else mask.astype(np.float64), scale,
               g.astype(np.float64))
    os.environ["PDN_ATTN"] = impl
    try:
        dev = "cuda:0"
        tq, tk, tv = (pdn.Tensor(t, dtype=np.float32, device=dev, requires_grad=True) for t in (q, k, v))
        tm = pdn.Tensor(mask, dtype=np.float32, device=dev) if mask is not None else None
        out = _fused.attention(tq, tk, tv, tm, float(scale))
        (out * pdn.Tensor(g, dtype=np.float32, device=dev)).sum().backward()
        got = (out.numpy(), tq.grad.get(), tk.grad.get(), tv.grad.get())
    finally:
        os.environ.pop("PDN_ATTN", None)
    for name, a, b in zip(("out", "dq", "dk", "dv"), got, ref):
        assert a.shape == b.shape, name
        assert np.isfinite(a).all(), name
        e = _err(a, b)
        assert e < 1e-4, f"{impl} {case} {name}: normwise rel err {e:.3e}"


ROW_CASES = [  # B, H, Lq, Lk, D, mask kind, cache length S (K/V are [:, :Lk] views of a [B, S, H, D] cache)
    (3, 6, 1, 130, 48, None, 256), (2, 4, 4, 37, 64, "causal", 37), (1, 2, 2, 300, 128, None, 512), (5, 3, 1, 9, 32, "pad", 16),
    (420, 6, 1, 70, 48, None, 96), (2, 2, 16, 16, 24, "causal", 16), (400, 8, 1, 33, 20, "pad", 40)
]


@pytest.mark.parametrize("case", ROW_CASES)
def test_attention_rows_decode_shapes(case):
    """The coalesced float4 row kernel (k_attention_rows: decode and short prefill over strided KV-cache views) — forward
    against the fp64 chain, and backward (lse it wrote feeds k_attention_bwd) as well."""
    import pydynet_b200 as pdn
    from pydynet_b200.nn import _fused
    B, H, Lq, Lk, D, mk, S = case
    rng = np.random.default_rng(7 + Lk)
    q = rng.standard_normal((B, Lq, H, D)).astype(np.float32)
    ck, cv = (rng.standard_normal((B, S, H, D)).astype(np.float32) for _ in range(2))
    g = rng.standard_normal((B, Lq, H * D)).astype(np.float32)
    mask = None
    if mk == "causal":
        mask = np.triu(np.full((Lq, Lk), -np.inf, np.float32), k=1 + max(0, Lk - Lq))
    elif mk == "pad":
        mask = np.zeros((B, 1, 1, Lk), np.float32)
        mask[0, ..., Lk - 5:] = -np.inf
        mask[B - 1, ..., :2] = -np.inf
    scale = 1.0 / np.sqrt(D)
    k, v = ck[:, :Lk], cv[:, :Lk]
    ref = _ref(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), None if mask is None else mask.astype(np.float64), scale,
               g.astype(np.float64))
    os.environ["PDN_ATTN"] = "ffma"
    try:
        dev = "cuda:0"
        tq = pdn.Tensor(q, dtype=np.float32, device=dev, requires_grad=True)
        tck, tcv = (pdn.Tensor(t, dtype=np.float32, device=dev, requires_grad=True) for t in (ck, cv))
        tm = pdn.Tensor(mask, dtype=np.float32, device=dev) if mask is not None else None
        out = _fused.attention(tq, tck[:, :Lk], tcv[:, :Lk], tm, float(scale))
        (out * pdn.Tensor(g, dtype=np.float32, device=dev)).sum().backward()
        got = (out.numpy(), tq.grad.get(), tck.grad.get()[:, :Lk], tcv.grad.get()[:, :Lk])
    finally:
        os.environ.pop("PDN_ATTN", None)
    for name, a, b in zip(("out", "dq", "dk", "dv"), got, ref):
        assert a.shape == b.shape, name
        assert np.isfinite(a).all(), name
        e = _err(a, b)
        assert e < 1e-4, f"rows {case} {name}: normwise rel err {e:.3e}"
