import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
os.environ.setdefault("X", "1")
import pydynet_b200 as pdn
from test_baseline_sizes import *
from test_baseline_sizes import _ref_class
def run(dev):
    from workloads.encoder import Transformer as St
    Transformer = _ref_class("examples/pydynet/transformer.py", "Transformer", (52, 192), lambda: St)
    np.random.seed(0)
    net = Transformer(512, 1, 8, 3, 0.05, 8192, 128); net.word_embedding.reset_parameters(); net.to(dev)
    rng = np.random.default_rng(2)
    X, y = rng.integers(1, 8192, (8, 128)), rng.choice([-1, 1], 8).astype(f32)
    net.train()
    out = net(T(X, dev), None)
    loss = pdn.log(1 + pdn.exp(-T(y, dev) * pdn.squeeze(out))).mean()
    loss.backward()
    return {k: np.asarray(p.grad.get() if hasattr(p.grad, "get") else p.grad) for k, p in net._parameters.items() if p.requires_grad}, out.numpy()
gc, oc = run("cpu"); gg, og = run("cuda:0")
rel = lambda a, b: np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30)
print("out gpu-vs-golden", rel(og, G["c4.out0"]), "cpu-vs-golden", rel(oc, G["c4.out0"]))
for k in gc:
    print(f"{k:45s} gpu-vs-ref {rel(thin(gg[k]), G['c4.g.'+k]):.2e}  ourcpu-vs-ref {rel(thin(gc[k]), G['c4.g.'+k]):.2e}  gpu-vs-ourcpu {rel(gg[k], gc[k]):.2e}")
