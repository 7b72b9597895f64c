"""1-D family golden vectors from the UNMODIFIED reference: conv1d / max_pool1d / avg_pool1d forward AND
backward (reference pydynet/nn/functional.py:61-191). The reference's conv1d expression `(col @ kernel.transpose(1, 2, 0)).sum(1)`
only broadcasts when the number of output positions equals the kernel size, so the conv1d cases are chosen that way (the only inputs
for which the reference defines a result); the pooling functions reduce the LAST axis of the (N, C, k, n_out) window tensor, i.e. they
return (N, C, k) — pinned as the reference computes it. Run in the build container: python tests/golden/make_golden_1d.py"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
import pydynet as pdn  # noqa: E402
import pydynet.nn.functional as F  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
rng = np.random.default_rng(11)
d = {}


def T(a, rg=False):
    return pdn.Tensor(np.asarray(a), dtype=np.asarray(a).dtype, requires_grad=rg)


# conv1d: (length, k, stride, pad) with n_out == k
for i, (Lx, k, stride, pad) in enumerate([(5, 3, 1, 0), (9, 4, 2, 1), (7, 5, 1, 1)]):
    N, C, O = 3, 4, 4 if i != 1 else 4
    x, w = rng.standard_normal((N, C, Lx)).astype(f32), rng.standard_normal((O, C, k)).astype(f32)
    n_out = (Lx + 2 * pad - k) // stride + 1
    assert n_out == k, (n_out, k)
    tx, tw = T(x, True), T(w, True)
    out = F.conv1d(tx, tw, pad, stride)
    g = rng.standard_normal(out.shape).astype(f32)
    (out * T(g)).sum().backward()
    d[f"conv1d{i}.cfg"] = np.array([stride, pad])
    d[f"conv1d{i}.x"], d[f"conv1d{i}.w"], d[f"conv1d{i}.g"] = x, w, g
    d[f"conv1d{i}.out"], d[f"conv1d{i}.dx"], d[f"conv1d{i}.dw"] = out.data.copy(), np.array(tx.grad), np.array(tw.grad)

for i, (Lx, k, stride, pad) in enumerate([(8, 2, 2, 0), (9, 3, 1, 1), (10, 3, 2, 2)]):
    x = rng.standard_normal((2, 3, Lx)).astype(f32)
    for nm, fn in (("max", F.max_pool1d), ("avg", F.avg_pool1d)):
        tx = T(x, True)
        out = fn(tx, k, stride, pad)
        g = rng.standard_normal(out.shape).astype(f32)
        (out * T(g)).sum().backward()
        d[f"pool1d{i}.cfg"] = np.array([k, stride, pad])
        d[f"pool1d{i}.x"] = x
        d[f"pool1d{i}.{nm}.g"], d[f"pool1d{i}.{nm}.out"], d[f"pool1d{i}.{nm}.dx"] = g, out.data.copy(), np.array(tx.grad)

np.savez_compressed(os.path.join(HERE, "family_1d.npz"), **d)
print("family_1d.npz:", len(d), "arrays")
