import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn, pydynet_b200.nn as nn, pydynet_b200.nn.functional as F
from oracle import pdn_oracle as O
f32 = np.float32
rng = np.random.default_rng(0)
worst = 0
for (N, C, H, W, Oc, k, pad) in [(5, 1, 28, 28, 20, 3, 1), (3, 3, 17, 13, 7, 3, 0), (4, 1, 12, 9, 6, 5, 2), (2, 2, 8, 8, 64, 3, 2), (300, 1, 28, 28, 20, 3, 1)]:
    x = rng.standard_normal((N, C, H, W)).astype(f32); w = rng.standard_normal((Oc, C, k, k)).astype(f32) * .3; b = rng.standard_normal(Oc).astype(f32)
    tx, tw, tb = (pdn.Tensor(a, dtype=f32, requires_grad=True, device="cuda:0") for a in (x, w, b))
    conv = nn.Conv2d(C, Oc, k, 1, pad, dtype=f32).to("cuda:0")
    with conv.weight.device:
        conv.weight.data[...] = w
        conv.bias.data[...] = b.reshape(1, -1, 1, 1)
    tw, tb = conv.weight, conv.bias
    y = conv(tx)
    gyv = rng.standard_normal(y.shape).astype(f32)
    (y * pdn.Tensor(gyv, dtype=f32, device="cuda:0")).sum().backward()
    # fp64 reference
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    oh, ow = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    yr = np.zeros((N, Oc, oh, ow)); dwr = np.zeros((Oc, C, k, k))
    for ky in range(k):
        for kx in range(k):
            patch = xp[:, :, ky:ky + oh, kx:kx + ow]
            yr += np.einsum("nchw,oc->nohw", patch, w[:, :, ky, kx].astype(np.float64))
            dwr[:, :, ky, kx] = np.einsum("nohw,nchw->oc", gyv.astype(np.float64), patch)
    yr += b[None, :, None, None]
    e = lambda a, r: np.linalg.norm(a - r) / np.linalg.norm(r)
    errs = (e(y.numpy(), yr), e(tw.grad.get() if hasattr(tw.grad, "get") else tw.grad, dwr), e(np.asarray(tb.grad.get() if hasattr(tb.grad, "get") else tb.grad).ravel(), gyv.sum((0, 2, 3))))
    print((N, C, H, W, Oc, k, pad), ["%.2e" % v for v in errs])
    worst = max(worst, *errs)
assert worst < 5e-6, worst
print("thin conv ok")
