"""PDN_TC_TRACE timelines + per-pass CUDA-event times of the fused attention at the bench_all shape (B128 H8 S512 hd64)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import pydynet_b200 as pdn
from pydynet_b200.nn import _fused
f32 = np.float32
B, H, S, D = (int(os.environ.get(k, d)) for k, d in (("B", 128), ("H", 8), ("S", 512), ("D", 64)))
rng = np.random.default_rng(0)
q, k, v = (pdn.Tensor(rng.standard_normal((B, S, H, D)).astype(f32), dtype=f32, requires_grad=True, device="cuda:0") for _ in range(3))


def att():
    for t in (q, k, v):
        t.zero_grad()
        t.data.buf.version += 1
    _fused.attention(q, k, v, None, 1.0 / D**.5).sum().backward()


for _ in range(2):
    att()
pdn.cuda.synchronize()
if not os.environ.get("PDN_TC_TRACE"):
    import time
    t0 = time.perf_counter()
    for _ in range(10):
        att()
    pdn.cuda.synchronize()
    print(f"attention B{B} H{H} S{S} D{D} fwd+bwd: {(time.perf_counter() - t0) * 100:.3f} ms")
